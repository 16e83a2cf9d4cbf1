"""Replay of REAL reference outputs, when somebody has produced them.

oracle/julia/gen_reference_fixtures.jl runs the unmodified Julia reference (rollout!, resolve, gradient(T) with the recorded rand(dim)
stream) and writes tests/golden/julia_*.npz. Julia is not installed in the build environment, so those files do not exist yet and
these tests SKIP; once they are committed, everything on the path except Optim.jl's IPNewton iterates is pinned to the reference:
the fixtures are replayed TEACHER-FORCED on the reference's own x-path through the CPU oracle (CPU test) and through the CUDA path
(GPU test). Tolerances: 1e-8 relative on draws / values (north_star), 1e-6 on gradients relative to the largest component.
"""
import glob
import os

import numpy as np
import pytest

from conftest import relerr

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "julia_*.npz")))
needs_fixtures = pytest.mark.skipif(not FIXTURES, reason="no tests/golden/julia_*.npz: run oracle/julia/gen_reference_fixtures.jl with Julia + the reference checkout")


def _check(f, r):
    assert relerr(r["ys"], f["ys"]) < 1e-8 and relerr(r["values"], f["values"]) < 1e-8
    assert relerr(r["gys"], f["gys"], floor=max(1.0, float(np.abs(f["gys"]).max()))) < 1e-8
    assert np.array_equal(r["best_index"], f["best_index"]) and np.array_equal(r["grad_case"], f["grad_case"])
    gscale = np.maximum(np.abs(f["grad_x"]).max(axis=0, keepdims=True), 1e-6)
    assert np.max(np.abs(r["grad_x"] - f["grad_x"]) / gscale) < 1e-6
    tscale = np.maximum(np.abs(f["grad_theta"]).max(axis=0, keepdims=True), 1e-6)
    assert np.max(np.abs(r["grad_theta"] - f["grad_theta"]) / tscale) < 1e-6


@needs_fixtures
@pytest.mark.parametrize("path", FIXTURES or ["<none>"])
def test_oracle_reproduces_reference_fixture(orc, path):
    f = np.load(path)
    h = int(f["h"])
    P = orc.OracleProblem(f["X"], f["L"], f["y"], f["c"], f["x0"], f["lbs"], f["ubs"], f["rn"], f["starts"], h=h, kernel="matern52", ktheta=(float(f["ell"]),),
                          rule="EI", theta=f["theta"], sigma_n2=float(f["sigma_n2"]), fmini=float(f["fmini"]), mode=1, dual_dirs=f["dual_dirs"],
                          x_forced=np.asfortranarray(f["xs"][:, 1:, :]))
    _check(f, P.rollout())


@needs_fixtures
@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES or ["<none>"])
def test_cuda_path_reproduces_reference_fixture(pkg, path):
    f = np.load(path)
    h, N = int(f["h"]), f["X"].shape[1]
    d, M = f["X"].shape[0], f["values"].shape[0]
    sur = pkg.Surrogate(pkg.Matern52([float(f["ell"])]), f["X"], f["y"], capacity=N + h + 1, decision_rule=pkg.EI(), σn2=float(f["sigma_n2"]))
    sur.L[:N, :N] = f["L"]; sur.c[:N] = f["c"]  # the reference's own factor and coefficients
    eng = pkg.RolloutEngine(0)
    try:
        eng.set_surrogate(pkg.FantasySurrogate(sur, h)); eng.set_normals(f["rn"]); eng.set_starts(f["starts"])
        r = dict(values=np.zeros(M), grad_x=np.zeros((d, M), order="F"), grad_theta=np.zeros((len(f["theta"]), M), order="F"),
                 best_index=np.zeros(M, np.int32), grad_case=np.zeros(M, np.int32))
        eng.rollout(f["x0"], f["theta"], f["lbs"], f["ubs"], h, float(f["fmini"]), r["values"], r["grad_x"], r["grad_theta"], dual_dirs=f["dual_dirs"],
                    x_forced=np.asfortranarray(f["xs"][:, 1:, :]), best_index=r["best_index"], grad_case=r["grad_case"])
        r.update(eng.tape(h))
    finally:
        eng.close()
    _check(f, r)


def test_fixture_generator_is_present_and_cites_the_reference():
    """The generator script is the committed road from 'parity unpinned' to pinned; it must keep citing what it records."""
    src = open(os.path.join(os.path.dirname(HERE), "oracle", "julia", "gen_reference_fixtures.jl")).read()
    for needle in ("rollout!(", "gradient(T)", "rollout.jl:133", "StochasticObservable(", "julia_", "ROLLOUT_BO_REFERENCE_DIR"):
        assert needle in src
