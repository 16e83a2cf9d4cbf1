"""C4-style stochastic gradient ascent loop on the rollout estimator (rollout_bayesopt.jl:87-129 shape: M samples, Adam, ESWAVS
early stopping) with the surrogate / normals / starts resident on the device; prints seconds per optimizer iteration."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as g

def main(name="C4", iters=20):
    pkg = g.load_package()
    wl = pkg.problems.make_workload(name)
    sur = wl.surrogate()
    fs = pkg.FantasySurrogate(sur, wl.h)
    T = pkg.Trajectory(sur, fs, start=wl.x0, hypers=wl.theta, horizon=wl.h)
    tp = pkg.TrajectoryParameters(wl.x0, wl.theta, wl.h, wl.M, True, wl.lbs, wl.ubs)
    es = pkg.ExperimentSetup(tp, wl.S)
    opt = pkg.Adam(η=0.01)
    t0 = time.perf_counter()
    x, history = pkg.stochastic_solve(opt, T, tp, es, wl.x0, max_iterations=iters, use_eswavs=False)
    dt = time.perf_counter() - t0
    n = len(history)
    print(f"{name}: d={wl.d} n={sur.observed} h={wl.h} M={wl.M} starts={wl.S}+2: {n} optimizer iterations in {dt:.2f} s = {dt / n * 1e3:.1f} ms per iteration "
          f"({wl.M * n / dt:.0f} trajectories/s incl. host loop); estimate {history[0][1]:.6f} -> {history[-1][1]:.6f}")

if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "C4", int(sys.argv[2]) if len(sys.argv) > 2 else 20)
