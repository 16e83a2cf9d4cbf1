"""Quick probe of one workload on the GPU: throughput, evaluation counts and (optionally) an oracle comparison.
   python scripts/probe.py C5 296 [--oracle 8] [--large-n] [--slots W] [--value-only]"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("name"); ap.add_argument("M", type=int)
    ap.add_argument("--oracle", type=int, default=0); ap.add_argument("--large-n", action="store_true")
    ap.add_argument("--slots", type=int, default=None); ap.add_argument("--value-only", action="store_true")
    ap.add_argument("--reps", type=int, default=2); ap.add_argument("--rs-cap", type=int, default=0)
    a = ap.parse_args()
    pkg = g.load_package()
    wl = pkg.problems.make_workload(a.name, M=a.M)
    t0 = time.time(); sur = wl.surrogate(); t_fit = time.time() - t0
    eng = pkg.RolloutEngine(0)
    if a.large_n or a.slots is not None:
        eng.set_tuning(large_n=a.large_n or None, large_n_slots=a.slots)
    if a.rs_cap:
        eng.handle.check(eng.lib.rbo_set_tuning(eng.handle.h, 4, a.rs_cap))
    t0 = time.time(); eng.set_surrogate(pkg.FantasySurrogate(sur, wl.h)); t_set = time.time() - t0
    eng.generate_normals(a.M, wl.h + 1)
    starts = pkg.generate_initial_guesses(wl.S, wl.lbs, wl.ubs)
    eng.set_starts(starts)
    dd = np.asfortranarray(np.random.default_rng(7).random((wl.d, wl.h, a.M)))
    vals = np.zeros(a.M); gx = None if a.value_only else np.zeros((wl.d, a.M), order="F"); gt = None if a.value_only else np.zeros((1, a.M), order="F")
    st = np.zeros(a.M, np.int32)
    fmini = float(np.min(sur.y))
    for _ in range(a.reps):
        s = eng.rollout(wl.x0, wl.theta, wl.lbs, wl.ubs, wl.h, fmini, vals, gx, gt, dual_dirs=None if a.value_only else dd, status=st)
    print(f"{a.name} M={a.M} d={wl.d} n={wl.N} h={wl.h} S={wl.S}+2: fit {t_fit:.2f}s set_surrogate {t_set:.3f}s kernel_ms={s.kernel_ms:.2f} traj/s={a.M / s.kernel_ms * 1e3:.1f} "
          f"evals/traj={s.n_evals / a.M:.1f} failed={s.n_failed} mean={s.mean:.10g} flops={s.flops:.4g} ({s.flops / s.kernel_ms / 1e9:.2f} TF/s alg, {s.flops_executed / s.kernel_ms / 1e9:.2f} TF/s exec)", flush=True)
    if a.oracle:
        from oracle import oracle as orc
        m = min(a.oracle, a.M)
        rn = eng.get_normals(wl.h + 1)
        rn_s = np.asfortranarray(rn[:m])
        import copy
        wl2 = copy.copy(wl); wl2.M = m
        P = g._oracle_problem(orc, wl2, sur, rn_s, starts, 0 if a.value_only else 1, dual_dirs=np.asfortranarray(dd[:, :, :m]))
        t0 = time.time(); r = P.rollout(tape=True); dt = time.time() - t0
        ev = np.abs(vals[:m] - r["values"]) / np.maximum(1.0, np.abs(r["values"]))
        print(f"oracle {m} trajectories in {dt:.1f}s ({m / dt:.2f} traj/s): free-running max rel err values {ev.max():.3e}; within 1e-8: {(ev < 1e-8).mean():.3f}")
        if not a.value_only:
            sc = np.maximum(1e-6, np.abs(r["grad_x"]).max(axis=0))
            eg = np.abs(gx[:, :m] - r["grad_x"]).max(axis=0) / sc
            print(f"  grad_x rel err (per trajectory, scaled by its largest component): max {eg.max():.3e}; within 1e-5: {(eg < 1e-5).mean():.3f}; cases {np.bincount(r['grad_case'], minlength=4)}")
    eng.close()


if __name__ == "__main__":
    main()
