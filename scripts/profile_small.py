"""Small single-launch driver for ncu: C3 shapes, a few trajectories per SM."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import __graft_entry__ as g

def main(name="C3", M=296, reps=2):
    pkg = g.load_package()
    wl = pkg.problems.make_workload(name, M=M)
    sur = wl.surrogate()
    eng = pkg.RolloutEngine(0)
    eng.set_surrogate(pkg.FantasySurrogate(sur, wl.h))
    eng.generate_normals(M, wl.h + 1)
    eng.set_starts(pkg.generate_initial_guesses(wl.S, wl.lbs, wl.ubs))
    dd = np.asfortranarray(np.random.default_rng(7).random((wl.d, wl.h, M)))
    vals, gx, gt = np.zeros(M), np.zeros((wl.d, M), order="F"), np.zeros((1, M), order="F")
    for _ in range(reps):
        s = eng.rollout(wl.x0, wl.theta, wl.lbs, wl.ubs, wl.h, float(np.min(sur.y)), vals, gx, gt, dual_dirs=dd)
    print(f"{name} M={M}: kernel_ms={s.kernel_ms:.2f} traj/s={M / s.kernel_ms * 1e3:.1f} evals/traj={s.n_evals / M:.1f} "
          f"TF/s alg={s.flops / s.kernel_ms / 1e9:.3f} exec={s.flops_executed / s.kernel_ms / 1e9:.3f} mean={s.mean:.6f}")
    eng.close()

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "C3", int(sys.argv[2]) if len(sys.argv) > 2 else 296)
