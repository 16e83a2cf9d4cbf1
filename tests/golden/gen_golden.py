"""Generates the fixtures of tests/golden/ (run from the repository root: python tests/golden/gen_golden.py).

WHAT THESE ARE: input/output vectors of the hot path produced by oracle/rbo_oracle.cpp (the C++ restatement of the
reference) on small seeded problems, each cross-checked here against the independent numpy restatement
(oracle/py_restatement.py) before it is written. They are NOT outputs of the Julia reference -- Julia is not available in
this environment and the reference ships no vectors for this path -- so parity stays "unpinned";
what the fixtures pin is the oracle itself (tests/test_oracle_cpu.py::test_oracle_reproduces_golden) and, without any
oracle call at test time, the CUDA path (tests/test_gpu_parity.py::test_golden_*).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from oracle import py_restatement as pr  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def relerr(a, b, floor=1.0):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(floor, np.abs(b))))


def case(pkg, name, M, N, h, S, seed, gh_nodes_n=0):
    wl = pkg.problems.make_workload(name, M=M, N=N, h=h, S=S, seed=seed)
    sur = wl.surrogate()
    starts = orc.generate_initial_guesses(S, wl.lbs, wl.ubs)
    dd = np.asfortranarray(np.random.default_rng(seed).random((wl.d, max(h, 1), M)))
    kw = {}
    if gh_nodes_n:
        nodes, weights = pkg.gausshermite(gh_nodes_n)
        idx = np.asarray(pkg.generate_indices(gh_nodes_n, h + 1)) - 1
        assert len(idx) == M
        kw = dict(gh_nodes=np.asfortranarray(nodes[idx].T), gh_weights=np.asfortranarray(weights[idx].T))
        rn = np.zeros((M, wl.d + 1, h + 1), order="F")
    else:
        rn = orc.gen_low_discrepancy_sequence(M, wl.d, h + 1)
    P = g._oracle_problem(orc, wl, sur, rn, starts, 1, dual_dirs=dd, **kw)
    r = P.rollout()
    assert np.all(r["status"] == 0)
    fmini = float(np.min(sur.y))
    if not gh_nodes_n:  # cross-check with the numpy restatement (Monte-Carlo observable only)
        cases = set()
        for m in range(M):
            fs, obs, grads = pr.rollout_teacher_forced(wl.X, wl.y, wl.ell, wl.sigma_n2, h, wl.x0, wl.theta, rn[m], r["xs"][:, 1:, m])
            assert relerr(obs, r["ys"][:, m]) < 1e-9 and relerr(grads, r["gys"][:, :, m]) < 1e-8
            gx, gth, c, t = pr.trajectory_gradient(fs, obs, grads, wl.theta, fmini, dd[:, :, m])
            assert c == r["grad_case"][m] and t == r["best_index"][m]
            assert relerr(gx, r["grad_x"][:, m], floor=max(1e-6, np.abs(gx).max())) < 1e-6
            cases.add(c)
        print(f"  numpy restatement agrees on {M} trajectories, gradient cases {sorted(cases)}")
    N_ = sur.observed
    out = dict(X=sur.X[:, :N_], L=sur.L[:N_, :N_], y=sur.y[:N_], c=sur.c[:N_], ell=wl.ell, sigma_n2=wl.sigma_n2, x0=wl.x0, theta=wl.theta,
               lbs=wl.lbs, ubs=wl.ubs, h=h, fmini=fmini, rn=rn, starts=starts, dual_dirs=dd,
               xs=r["xs"], ys=r["ys"], gys=r["gys"], values=r["values"], grad_x=r["grad_x"], grad_theta=r["grad_theta"],
               best_index=r["best_index"], grad_case=r["grad_case"], alphas=r["alphas"])
    out.update(kw)
    return out


def main():
    g.build() if not os.path.exists(os.path.join(ROOT, "oracle", "librbo_oracle.so")) else None
    pkg = g.load_package()
    print("mc_hartmann6: Matern52 + EI, d=6, N=14, h=2, 16 trajectories, 4+2 starts")
    np.savez_compressed(os.path.join(HERE, "mc_hartmann6.npz"), **case(pkg, "C2", 16, 14, 2, 4, 3))
    print("mc_gp2d: d=2, N=12, h=3, 24 trajectories (case-3 gradients survive the det test)")
    np.savez_compressed(os.path.join(HERE, "mc_gp2d.npz"), **case(pkg, "GP:2:0.25", 24, 12, 3, 4, 5))
    print("ghq_hartmann6: Gauss-Hermite observable, 3 nodes, depth 3 (27 trajectories)")
    np.savez_compressed(os.path.join(HERE, "ghq_hartmann6.npz"), **case(pkg, "C2", 27, 14, 2, 4, 3, gh_nodes_n=3))


if __name__ == "__main__":
    main()
