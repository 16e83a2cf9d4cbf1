"""Restarts of the stochastic-ascent loop: n_x0 serial estimator calls vs one rbo_rollout_batch launch (reference-sized M)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g

def main(name="C2", M=64, B=8):
    pkg = g.load_package()
    wl = pkg.problems.make_workload(name, M=M)
    sur = wl.surrogate()
    eng = pkg.RolloutEngine(0)
    eng.set_surrogate(pkg.FantasySurrogate(sur, wl.h))
    eng.generate_normals(M, wl.h + 1)
    eng.set_starts(pkg.generate_initial_guesses(wl.S, wl.lbs, wl.ubs))
    dd = np.asfortranarray(np.random.default_rng(7).random((wl.d, wl.h, M)))
    x0s = np.asfortranarray(wl.lbs[:, None] + (wl.ubs - wl.lbs)[:, None] * np.random.default_rng(1).random((wl.d, B)))
    fmini = float(np.min(sur.y))
    vb, gxb, gtb = np.zeros((M, B), order="F"), np.zeros((wl.d, M, B), order="F"), np.zeros((1, M, B), order="F")
    v, gx, gt = np.zeros(M), np.zeros((wl.d, M), order="F"), np.zeros((1, M), order="F")
    for rep in range(2):
        t0 = time.perf_counter()
        for b in range(B):
            eng.rollout(x0s[:, b], wl.theta, wl.lbs, wl.ubs, wl.h, fmini, v, gx, gt, dual_dirs=dd)
        t1 = time.perf_counter()
        eng.rollout_batch(x0s, wl.theta, wl.lbs, wl.ubs, wl.h, fmini, vb, gxb, gtb, dual_dirs=dd)
        t2 = time.perf_counter()
    print(f"{name} d={wl.d} n={sur.observed} h={wl.h} M={M} starts={wl.S}+2, {B} starting points: serial {1e3 * (t1 - t0):.1f} ms, one batch launch {1e3 * (t2 - t1):.1f} ms "
          f"({(t1 - t0) / (t2 - t1):.2f}x)")
    eng.close()

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "C2", int(sys.argv[2]) if len(sys.argv) > 2 else 64, int(sys.argv[3]) if len(sys.argv) > 3 else 8)
