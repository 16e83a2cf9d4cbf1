"""Independent numpy/scipy restatement of the reference arithmetic, used ONLY to cross-check the C++
oracle (tests/test_oracle_*.py). TEST INFRASTRUCTURE; PARITY UNPINNED (no Julia available).

It is deliberately written differently from rbo_oracle.cpp: dense matrices, LAPACK triangular solves
(scipy, the same kind of backend Julia's LinearAlgebra uses), derivatives of the kernel by closed form
checked against complex-step / finite differences in the tests, and the adjoint loops transcribed
one-to-one from rollout.jl:114-277 with the perturbation surrogates of rbs.jl:633-764 built densely
(rbf.jl:210-262), exactly as the reference does.

Only Matern-5/2 + EI are restated here (the configuration every BASELINE config uses).
"""
import numpy as np
from scipy.linalg import cholesky, solve_triangular
from scipy.special import erfc


def psi52(rho, ell):  # rbf.jl:60-68
    s = np.sqrt(5.0) * rho / ell
    return (1 + s * (1 + s / 3.0)) * np.exp(-s)


def dpsi52(rho, ell):
    c2 = 5.0 / ell**2
    s = np.sqrt(5.0) * rho / ell
    return -(c2 * rho / 3.0) * (1 + s) * np.exp(-s)


def d2psi52(rho, ell):
    c2 = 5.0 / ell**2
    s = np.sqrt(5.0) * rho / ell
    return (c2 / 3.0) * (s * s - s - 1) * np.exp(-s)


def grad_k(r, ell):  # rbf.jl:127-134
    rho = np.linalg.norm(r)
    if rho == 0:
        return 0 * r
    return dpsi52(rho, ell) * r / rho


def hess_k(r, ell):  # rbf.jl:141-150
    p = np.linalg.norm(r)
    d = len(r)
    if p > 0:
        dp = r / p
        Dpr = dpsi52(p, ell) / p
        D2 = d2psi52(p, ell)
        return (D2 - Dpr) * np.outer(dp, dp) + Dpr * np.eye(d)
    return d2psi52(p, ell) * np.eye(d)


def KXX(X, ell, sn2):  # rbf.jl:161-178
    N = X.shape[1]
    K = np.zeros((N, N))
    for j in range(N):
        K[j, j] = psi52(0.0, ell)
        for i in range(j + 1, N):
            K[i, j] = K[j, i] = psi52(np.linalg.norm(X[:, i] - X[:, j]), ell)
    return K + sn2 * np.eye(N)


def ei_partials(mu, sigma, theta1, fstar, sigma_tol=1e-8):
    """decision_rules.jl:84-99 and the AD partials of decision_rules.jl:23-34 (closed forms)."""
    if sigma < sigma_tol:
        return dict(g=0.0, g_mu=0.0, g_sig=0.0, g_mumu=0.0, g_sigsig=0.0, g_muth=0.0, g_sigth=0.0)
    imp = fstar - mu - theta1
    z = imp / sigma
    Phi = 0.5 * erfc(-z / np.sqrt(2.0))
    phi = np.exp(-0.5 * z * z) / np.sqrt(2 * np.pi)
    return dict(g=imp * Phi + sigma * phi, g_mu=-Phi, g_sig=phi, g_mumu=phi / sigma, g_sigsig=z * z * phi / sigma,
                g_muth=phi / sigma, g_sigth=z * phi / sigma)


class Fantasy:
    """rbs.jl:320-480."""

    def __init__(self, X, y, ell, sn2, h):
        self.d, self.N = X.shape
        self.ell, self.sn2, self.h = ell, sn2, h
        cap = self.N + h + 1
        self.X = np.zeros((self.d, cap))
        self.X[:, : self.N] = X
        self.y = np.zeros(cap)
        self.y[: self.N] = y
        K = KXX(X, ell, sn2)
        self.L = np.zeros((cap, cap))
        self.L[: self.N, : self.N] = cholesky(K, lower=True)
        L = self.L[: self.N, : self.N]
        self.cs = [solve_triangular(L.T, solve_triangular(L, y, lower=True), lower=False)]
        self.nf = 0

    def condition(self, x, yv):  # rbs.jl:431-441
        n = self.N + self.nf + 1
        self.X[:, n - 1] = x
        self.y[n - 1] = yv
        self.nf += 1
        kx = np.array([psi52(np.linalg.norm(x - self.X[:, j]), self.ell) for j in range(n - 1)])
        L = self.L[: n - 1, : n - 1]
        L21 = solve_triangular(L, kx, lower=True)
        self.L[n - 1, : n - 1] = L21
        self.L[n - 1, n - 1] = np.sqrt(psi52(0.0, self.ell) + self.sn2 - L21 @ L21)
        Ln = self.L[:n, :n]
        self.cs.append(solve_triangular(Ln.T, solve_triangular(Ln, self.y[:n], lower=True), lower=False))

    def eval(self, x, theta, fantasy_index):  # rbs.jl:482-581
        n = self.N + fantasy_index + 1
        X, L, c, y = self.X[:, :n], self.L[:n, :n], self.cs[fantasy_index + 1], self.y[:n]
        d, ell = self.d, self.ell
        s = {"x": np.array(x, float), "n": n, "c": c, "fantasy_index": fantasy_index}
        kx = np.array([psi52(np.linalg.norm(x - X[:, j]), ell) for j in range(n)])
        dkx = np.stack([grad_k(x - X[:, j], ell) for j in range(n)], axis=1)  # d x n
        Ksolve = lambda B: solve_triangular(L.T, solve_triangular(L, B, lower=True), lower=False)
        mu = kx @ c
        dmu = dkx @ c
        Hk = [hess_k(x - X[:, j], ell) for j in range(n)]
        Hmu = sum(c[j] * Hk[j] for j in range(n))
        w = Ksolve(kx)
        Dw = Ksolve(dkx.T)  # n x d
        sigma = np.sqrt(psi52(0.0, ell) - kx @ w)
        dsig = -(dkx @ w) / sigma
        Hsig = (-np.outer(dsig, dsig) - dkx @ Dw - sum(w[j] * Hk[j] for j in range(n))) / sigma
        fstar = y.min()
        g = ei_partials(mu, sigma, theta[0], fstar)
        dal = g["g_mu"] * dmu + g["g_sig"] * dsig
        Hal = g["g_mumu"] * np.outer(dmu, dmu) + g["g_mu"] * Hmu + g["g_sigsig"] * np.outer(dsig, dsig) + g["g_sig"] * Hsig
        s.update(kx=kx, dkx=dkx, mu=mu, dmu=dmu, Hmu=Hmu, w=w, Dw=Dw, sigma=sigma, dsig=dsig, Hsig=Hsig, fstar=fstar,
                 g=g, alpha=g["g"], dal=dal, Hal=Hal, d2a_dxdth=dmu * g["g_muth"] + dsig * g["g_sigth"])
        return s

    def draw(self, x, theta, fantasy_index, z):  # rbs.jl:588-611 with dsigma rbs.jl:530-539
        s = self.eval(x, theta, fantasy_index)
        n, d, ell = s["n"], self.d, self.ell
        L = self.L[:n, :n]
        kxx = np.zeros((d + 1, d + 1))
        kxx[0, 0] = psi52(0.0, ell)
        kxx[1:, 1:] = -d2psi52(0.0, ell) * np.eye(d)
        kxX = np.vstack([s["kx"][None, :], s["dkx"]])
        Sg = kxx - kxX @ solve_triangular(L.T, solve_triangular(L, kxX.T, lower=True), lower=False)
        Sg = np.triu(Sg) + np.triu(Sg, 1).T  # Symmetric(A) takes the upper triangle
        Ls = cholesky(Sg, lower=True)
        return np.concatenate([[s["mu"]], s["dmu"]]) + Ls @ z

    def perturb(self, sx, theta, fantasy_step, sample_index, dx, spatial):  # rbs.jl:652-694 / 711-760
        n = self.N + fantasy_step + 1
        assert n == sx["n"]
        X, L, ell, d = self.X[:, :n], self.L[:n, :n], self.ell, self.d
        dX = np.zeros((d, n))
        dX[:, self.N + sample_index] = dx
        dK = np.zeros((n, n))
        for j in range(n):  # rbf.jl:210-228
            for i in range(j + 1, n):
                dK[i, j] = dK[j, i] = grad_k(X[:, i] - X[:, j], ell) @ (dX[:, i] - dX[:, j])
        c = self.cs[fantasy_step + 1]
        x = sx["x"]
        dc = -solve_triangular(L.T, solve_triangular(L, dK @ c, lower=True), lower=False)
        dkx = np.array([grad_k(x - X[:, j], ell) @ (-dX[:, j]) for j in range(n)])
        dgkx = np.stack([hess_k(x - X[:, j], ell) @ (-dX[:, j]) for j in range(n)], axis=1)
        dmu = dkx @ c + sx["kx"] @ dc
        dgmu = dgkx @ c + sx["dkx"] @ dc
        w, sigma = sx["w"], sx["sigma"]
        dsig = (-2 * dkx @ w + w @ (dK @ w)) / (2 * sigma)
        gh = ei_partials(dmu, dsig, theta[0], sx["fstar"])  # Q6: partials evaluated at the variations
        g = sx["g"]
        if spatial:
            dgsig = (sx["Dw"].T @ (dK @ w) - dgkx @ w - sx["Dw"].T @ dkx - dsig * sx["dsig"]) / sigma
            return g["g_mu"] * dgmu + g["g_sig"] * dgsig + gh["g_mu"] * sx["dmu"] + gh["g_sig"] * sx["dsig"]
        return g["g_mu"] * dgmu + gh["g_mu"] * sx["dmu"] + gh["g_sig"] * sx["dsig"]


def rollout_teacher_forced(X, y, ell, sn2, h, x0, theta, z, xpath):
    """rollout.jl:39-74 with the inner solves replaced by the given x_1..x_h. z is (d+1) x (h+1)."""
    fs = Fantasy(X, y, ell, sn2, h)
    obs, grads = np.zeros(h + 1), np.zeros((fs.d, h + 1))
    for step in range(h + 1):
        x = x0 if step == 0 else xpath[:, step - 1]
        dr = fs.draw(x, theta, step - 1, z[:, step])
        obs[step], grads[:, step] = dr[0], dr[1:]
        fs.condition(x, dr[0])
    return fs, obs, grads


def trajectory_gradient(fs, obs, grads, theta, fmini, dual_dirs, htol=1e-4):
    """rollout.jl:233-277. dual_dirs is d x h (column solve_index holds the rand(dim) of rollout.jl:133)."""
    d, N, h = fs.d, fs.N, fs.h
    yf = fs.y[N:N + h + 1]
    t = int(np.argmin(yf))
    fb = yf[t]
    if fmini <= fb:
        return np.zeros(d), np.zeros(1), 1, t
    if t == 0:
        return -grads[:, 0], np.zeros(1), 2, t
    rps = lambda j: fs.eval(fs.X[:, N + j], theta, j - 1)
    xbars = {j: np.zeros(d) for j in range(1, t + 1)}
    ybars = np.zeros(t + 2)
    ybars[t + 1] = 1.0
    I = np.eye(d)
    for j in range(t, 0, -1):
        sx = rps(j)
        if np.linalg.det(sx["Hal"]) < htol:
            xbars[j] = np.zeros(d)
        else:
            xd = -grads[:, j - 1] * ybars[j + 1]
            for i in range(j + 1, t + 1):
                sxi = rps(i)
                dri = np.zeros((d, d))
                for k in range(d):
                    dri[:, k] = fs.perturb(sxi, theta, i - 1, j, I[:, k], True)
                xd = xd - dri.T @ xbars[i]
            xbars[j] = np.linalg.solve(sx["Hal"].T, xd)
        sidx = j - 1
        yd = 0.0
        for i in range(sidx + 1, t + 1):
            sxi = rps(i)
            yd += fs.perturb(sxi, theta, i - 1, sidx, dual_dirs[:, sidx], False) @ xbars[i]
        ybars[j] = yd
    sx0 = rps(0)
    gx = sx0["dmu"] * ybars[1]
    gth = np.zeros(1)
    for j in range(1, t + 1):
        sxj = rps(j)
        G = np.zeros((d, d))
        for k in range(d):
            G[:, k] = fs.perturb(sxj, theta, j - 1, 0, I[:, k], True)
        gx = gx + G.T @ xbars[j]
        gth = gth + sxj["d2a_dxdth"] @ xbars[j]
    return -gx, -gth, 3, t
