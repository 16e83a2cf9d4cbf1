"""Per-phase cycle breakdown of the rollout kernel (needs csrc/librbo_timers.so built with -DRBO_PHASE_TIMERS)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.environ["RBO_LIB_PATH"] = os.path.join(ROOT, "rollout-bayesian-optimization_b200", "csrc", os.environ.get("RBO_TIMERS_LIB", "librbo_timers.so"))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g

NAMES = ["fill_columns", "reduce_pre", "fwd_solve", "reduce_post", "bwd_solve", "reduce_hess", "per-start logic (all slots: assemble + solver step, slowest warp)", "bookkeeping",
         "  logic: assemble (warp 0)", "round gap", "draw+condition", "adjoint", "wait_full(fwd,w0)", "wait_full(bwd,w0)", "  logic: solver step (warp 0)", "rounds"]

def main(name="C3", M=296, large_n=False):
    pkg = g.load_package()
    wl = pkg.problems.make_workload(name, M=M)
    sur = wl.surrogate()
    eng = pkg.RolloutEngine(0)
    if large_n:
        eng.set_tuning(large_n=True)
    eng.set_surrogate(pkg.FantasySurrogate(sur, wl.h))
    eng.generate_normals(M, wl.h + 1)
    eng.set_starts(pkg.generate_initial_guesses(wl.S, wl.lbs, wl.ubs))
    dd = np.asfortranarray(np.random.default_rng(7).random((wl.d, wl.h, M)))
    vals, gx, gt = np.zeros(M), np.zeros((wl.d, M), order="F"), np.zeros((1, M), order="F")
    out = (ctypes.c_ulonglong * 16)()
    sfx = "_largen" if (large_n or name == "C5") else ""
    fn = getattr(eng.lib, "rbo_debug_phase_cycles" + sfx)
    fn.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.c_int]
    s = eng.rollout(wl.x0, wl.theta, wl.lbs, wl.ubs, wl.h, float(np.min(sur.y)), vals, gx, gt, dual_dirs=dd)
    fn(eng.handle.h, out, 1)
    aux = (ctypes.c_ulonglong * 16)()
    fa = getattr(eng.lib, "rbo_debug_aux_cycles" + sfx)
    fa.argtypes = fn.argtypes
    fa(eng.handle.h, aux, 1)
    ft = getattr(eng.lib, "rbo_debug_tr_cycles" + sfx, None)
    trc = (ctypes.c_ulonglong * 16)()
    if ft is not None:
        ft.argtypes = fn.argtypes
        ft(eng.handle.h, trc, 1)
    s = eng.rollout(wl.x0, wl.theta, wl.lbs, wl.ubs, wl.h, float(np.min(sur.y)), vals, gx, gt, dual_dirs=dd)
    fn(eng.handle.h, out, 1)
    tot = sum(out[i] for i in range(12))
    rounds = out[15]
    print(f"{name} M={M}: kernel_ms={s.kernel_ms:.2f} traj/s={M / s.kernel_ms * 1e3:.1f} rounds/traj={rounds / M:.1f} evals/traj={s.n_evals / M:.1f} cycles/traj={tot / M:.3e}")
    for i in range(15):
        if NAMES[i] != "-":
            print(f"  {NAMES[i]:<60} {100 * out[i] / tot:5.1f}%   {out[i] / max(rounds, 1):9.0f} cycles/round")
    fa(eng.handle.h, aux, 0)
    tot_calls = sum(aux[i] for i in range(16))
    print("  per-start solver step (slot_logic_warp) duration histogram, bins of 4k cycles (last bin: >= 60k): " +
          " ".join(f"{100 * aux[i] / max(tot_calls, 1):.1f}%" for i in range(16)))
    if ft is not None:
        ft(eng.handle.h, trc, 0)
        calls = max(trc[0], 1)
        names = ["calls", "state machine before the step", "trust-region step", "after the step (projection, predicted decrease)", "  tr: load", "  tr: Householder",
                 "  tr: write-out + Gershgorin", "  tr: multisection probes", "  tr: tridiagonal solve (+ hard case)", "  tr: back-transformation"]
        print(f"  per-start logic by part (n <= 16 path), cycles per trust-region step over {trc[0]} steps:")
        for i in range(1, 10):
            print(f"    {names[i]:<52} {trc[i] / calls:9.0f}")
    eng.close()

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "C3", int(sys.argv[2]) if len(sys.argv) > 2 else 296, len(sys.argv) > 3 and sys.argv[3] == "--large-n")
