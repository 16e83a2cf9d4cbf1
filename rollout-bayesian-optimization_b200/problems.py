"""Synthetic workloads of BASELINE.json, seeded, FP64. Host-side data generators only."""
from dataclasses import dataclass

import numpy as np

from . import api


def branin(x):  # testfns.jl:136-152
    a, b, c, r, s, t = 1.0, 5.1 / (4 * np.pi**2), 5 / np.pi, 6.0, 10.0, 1 / (8 * np.pi)
    return a * (x[1] - b * x[0] ** 2 + c * x[0] - r) ** 2 + s * (1 - t) * np.cos(x[0]) + s


def hartmann6(x):  # testfns.jl:532-565
    al = np.array([1.0, 1.2, 3.0, 3.2])
    A = np.array([[10, 3, 17, 3.5, 1.7, 8], [0.05, 10, 17, 0.1, 8, 14], [3, 3.5, 1.7, 10, 17, 8], [17, 8, 0.05, 10, 0.1, 14]])
    P = 1e-4 * np.array([[1312, 1696, 5569, 124, 8283, 5886], [2329, 4135, 8307, 3736, 1004, 9991],
                         [2348, 1451, 3522, 2883, 3047, 6650], [4047, 8828, 8732, 5743, 1091, 381]])
    return -np.sum(al * np.exp(-np.sum(A * (x - P) ** 2, axis=1)))


@dataclass
class Workload:
    name: str
    d: int
    N: int
    h: int
    M: int
    S: int            # Sobol starts; the reference adds two corner starts (utils.jl:145-153)
    ell: float
    lbs: np.ndarray
    ubs: np.ndarray
    X: np.ndarray     # d x N
    y: np.ndarray
    x0: np.ndarray
    theta: np.ndarray
    with_grad: bool
    sigma_n2: float = 1e-6
    capacity_extra: int = 5

    def surrogate(self):
        return api.Surrogate(api.Matern52([self.ell]), self.X, self.y, capacity=self.N + self.capacity_extra, decision_rule=api.EI(), σn2=self.sigma_n2)


CONFIGS = {
    # name: (d, N, h, M, S, ell, with_grad)
    "C1": (2, 10, 1, 64, 8, 1.0, True),
    "C2": (6, 50, 3, 1024, 32, 0.5, False),
    "C3": (10, 200, 5, 16384, 8, 0.7, True),
    "C4": (6, 50, 4, 4096, 64, 0.5, True),
    "C5": (20, 1000, 2, 65536, 8, 1.5, True),
}


def make_workload(name, M=None, N=None, h=None, S=None, seed=1906):
    """BASELINE.json configs C1..C5. Observation sites are seeded uniform draws in the box (the reference's drivers
    use `rand`, nonmyopic_bayesopt.jl:206): using Sobol points for both the design and the inner-solve starts would
    make every start coincide with an observation. C3/C5 targets are a draw from the GP prior itself."""
    if name.startswith("GP:"):  # generic GP-prior workload "GP:<d>:<ell>" used by the tests
        _, ds, es = name.split(":")
        d, N0, h0, M0, S0, ell, with_grad = int(ds), 20, 2, 32, 4, float(es), True
    else:
        d, N0, h0, M0, S0, ell, with_grad = CONFIGS[name]
    N = N0 if N is None else N
    h = h0 if h is None else h
    M = M0 if M is None else M
    S = S0 if S is None else S
    rng = np.random.default_rng(seed)
    if name == "C1":
        lbs, ubs = np.array([-5.0, 0.0]), np.array([10.0, 15.0])
        X = lbs[:, None] + (ubs - lbs)[:, None] * rng.random((d, N))
        y = np.array([branin(X[:, j]) for j in range(N)])
    elif name in ("C2", "C4"):
        lbs, ubs = np.zeros(d), np.ones(d)
        X = rng.random((d, N))
        y = np.array([hartmann6(X[:, j]) for j in range(N)])
    else:
        lbs, ubs = np.zeros(d), np.ones(d)
        X = rng.random((d, N))
        K = api.eval_KXX(api.Matern52([ell]), X, 1e-6)
        y = np.linalg.cholesky(K) @ rng.standard_normal(N)
    x0 = 0.5 * (lbs + ubs)
    return Workload(name, d, N, h, M, S, ell, lbs, ubs, np.asfortranarray(X), y, x0, np.zeros(1), with_grad)
