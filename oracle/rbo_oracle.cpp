/*
 * rbo_oracle.cpp -- CPU restatement ("oracle") of the reference's rollout value + adjoint gradient path.
 *
 * TEST INFRASTRUCTURE ONLY (see rbo_oracle.h). PARITY UNPINNED: Julia is not available here and the
 * reference ships no golden vectors for this path; every function cites the reference file:line it restates.
 * Reference files are relative to the upstream repository root (rbs.jl = radial_basis_surrogates.jl,
 * rbf.jl = radial_basis_functions.jl).
 *
 * The arithmetic follows the reference, including its quirks. The only part that is
 * NOT a restatement is the inner box-constrained maximiser: the reference calls Optim.jl's IPNewton
 * (rbf_optim.jl:24-30; un-pinned third-party code, x_tol = f_tol = 1e-3). It is replaced here by a
 * regularised projected Newton iteration run to a tight tolerance; the CUDA path implements the same
 * algorithm independently, so free-running parity is parity with this restatement, not with Optim.jl.
 */
#include "rbo_oracle.h"
#include "sobol_joe_kuo.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const double kInvSqrt2 = 0.70710678118654752440;
const double kInvSqrt2Pi = 0.39894228040143267794;
const double kTwoPi = 6.283185307179586;  // Julia's 2*pi in Float64

// ----------------------------------------------------------------------------------------------
// Kernel scalar functions: rbf.jl:60-103 (psi) and their exact derivatives (rbf.jl:41-46 uses ForwardDiff)
// ----------------------------------------------------------------------------------------------
struct Kern {
  int id;
  double th[4];
};

inline void kern_eval(const Kern& k, double rho, double& psi, double& dpsi, double& d2psi) {
  switch (k.id) {
    case ORC_KERNEL_MATERN52: {  // rbf.jl:60-68
      double l = k.th[0], c = std::sqrt(5.0) / l, s = c * rho, e = std::exp(-s);
      psi = (1 + s * (1 + s / 3.0)) * e;
      dpsi = -(c * c * rho / 3.0) * (1 + s) * e;
      d2psi = (c * c / 3.0) * (s * s - s - 1) * e;
      break;
    }
    case ORC_KERNEL_MATERN32: {  // rbf.jl:70-78
      double l = k.th[0], c = std::sqrt(3.0) / l, s = c * rho, e = std::exp(-s);
      psi = (1 + s) * e;
      dpsi = -c * c * rho * e;
      d2psi = c * c * (s - 1) * e;
      break;
    }
    case ORC_KERNEL_MATERN12: {  // rbf.jl:80-88
      double l = k.th[0], s = rho / l, e = std::exp(-s);
      psi = e;
      dpsi = -e / l;
      d2psi = e / (l * l);
      break;
    }
    case ORC_KERNEL_SE: {  // rbf.jl:90-96
      double l = k.th[0], l2 = l * l;
      psi = std::exp(-rho * rho / (2 * l2));
      dpsi = -rho / l2 * psi;
      d2psi = (rho * rho / (l2 * l2) - 1.0 / l2) * psi;
      break;
    }
    default: {  // ORC_KERNEL_PERIODIC rbf.jl:98-103: exp(-2 sin(pi rho / th2)^2 / th1^2)
      double l = k.th[0], p = k.th[1], u = M_PI * rho / p, sn = std::sin(u);
      psi = std::exp(-2 * sn * sn / (l * l));
      double q = -(2 * M_PI / (p * l * l));
      dpsi = q * std::sin(2 * u) * psi;
      d2psi = q * (std::cos(2 * u) * (2 * M_PI / p) * psi + std::sin(2 * u) * dpsi);
      break;
    }
  }
}

// ----------------------------------------------------------------------------------------------
// Decision rules: decision_rules.jl:84-127. Partials are the closed forms of the ForwardDiff
// derivatives built at decision_rules.jl:23-34 (verified symbolically).
// ----------------------------------------------------------------------------------------------
inline double normcdf(double z) { return 0.5 * std::erfc(-z * kInvSqrt2); }
inline double normpdf(double z) { return std::exp(-0.5 * z * z) * kInvSqrt2Pi; }

struct GPart {
  double g, g_mu, g_sig, g_mumu, g_sigsig, g_th, g_thth, g_muth, g_sigth, g_musig;
};

inline GPart rule_eval(int rule, double sigma_tol, double mu, double sigma, double th1, double fstar) {
  GPart r;
  std::memset(&r, 0, sizeof(r));
  if (rule == ORC_RULE_LCB) {  // decision_rules.jl:117-127
    r.g = th1 * sigma - mu;
    r.g_mu = -1;
    r.g_sig = th1;
    r.g_th = sigma;
    r.g_sigth = 1;
    return r;
  }
  if (sigma < sigma_tol) return r;  // decision_rules.jl:87-89, 103-105: constant 0 => all partials 0
  double imp = fstar - mu - th1, z = imp / sigma, Phi = normcdf(z), phi = normpdf(z);
  if (rule == ORC_RULE_EI) {  // decision_rules.jl:84-99
    r.g = imp * Phi + sigma * phi;
    r.g_mu = -Phi;
    r.g_sig = phi;
    r.g_mumu = phi / sigma;
    r.g_sigsig = z * z * phi / sigma;
    r.g_th = -Phi;
    r.g_thth = phi / sigma;
    r.g_muth = phi / sigma;
    r.g_sigth = z * phi / sigma;
    r.g_musig = z * phi / sigma;
  } else {  // POI decision_rules.jl:101-115
    double s2 = sigma * sigma;
    r.g = Phi;
    r.g_mu = -phi / sigma;
    r.g_sig = -z * phi / sigma;
    r.g_mumu = -z * phi / s2;
    r.g_sigsig = (2 * z - z * z * z) * phi / s2;
    r.g_th = r.g_mu;
    r.g_thth = r.g_mumu;
    r.g_muth = r.g_mumu;
    r.g_sigth = (1 - z * z) * phi / s2;
    r.g_musig = (1 - z * z) * phi / s2;
  }
  return r;
}

// ----------------------------------------------------------------------------------------------
// Small dense helpers
// ----------------------------------------------------------------------------------------------
inline double dot(const double* a, const double* b, int n) {
  double s = 0;
#pragma omp simd reduction(+ : s)
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

// Row-major lower-triangular L (leading dimension ld). V holds nr right-hand sides, each contiguous of
// length n (V[r*ldv + i]).  Forward: V <- L^-1 V ; backward: V <- L^-T V.  Stand-ins for Julia's `L \ B`
// and `L' \ B` on LowerTriangular (LAPACK trtrs).
void fwd_solve(const double* L, int ld, int n, double* V, int ldv, int nr) {
  for (int i = 0; i < n; ++i) {
    const double* Li = L + (size_t)i * ld;
    for (int r = 0; r < nr; ++r) {
      double* v = V + (size_t)r * ldv;
      v[i] = (v[i] - dot(Li, v, i)) / Li[i];
    }
  }
}
void bwd_solve(const double* L, int ld, int n, double* V, int ldv, int nr) {
  for (int i = n - 1; i >= 0; --i) {
    const double* Li = L + (size_t)i * ld;
    for (int r = 0; r < nr; ++r) {
      double* v = V + (size_t)r * ldv;
      double wi = v[i] / Li[i];
      v[i] = wi;
#pragma omp simd
      for (int k = 0; k < i; ++k) v[k] -= Li[k] * wi;
    }
  }
}

// dense Cholesky (lower) of a symmetric matrix given by its UPPER triangle, as `cholesky(Symmetric(A))`
// (rbs.jl:536-537; Symmetric defaults to uplo = :U).  A is n x n row-major; result lower in Lc (row-major).
bool chol_from_upper(const double* A, int n, double* Lc) {
  for (int i = 0; i < n * n; ++i) Lc[i] = 0;
  for (int j = 0; j < n; ++j) {
    double s = A[j * n + j];
    for (int k = 0; k < j; ++k) s -= Lc[j * n + k] * Lc[j * n + k];
    if (!(s > 0)) return false;
    double ljj = std::sqrt(s);
    Lc[j * n + j] = ljj;
    for (int i = j + 1; i < n; ++i) {
      double t = A[j * n + i];  // upper triangle entry (j,i) stands for (i,j)
      for (int k = 0; k < j; ++k) t -= Lc[i * n + k] * Lc[j * n + k];
      Lc[i * n + j] = t / ljj;
    }
  }
  return true;
}

// LU with partial pivoting (Julia `det`, `\` on a dense square matrix: rollout.jl:159,188). A row-major, overwritten.
bool lu_factor(double* A, int n, int* piv, double* det) {
  double dt = 1;
  bool ok = true;
  for (int k = 0; k < n; ++k) {
    int p = k;
    double mx = std::fabs(A[k * n + k]);
    for (int i = k + 1; i < n; ++i)
      if (std::fabs(A[i * n + k]) > mx) { mx = std::fabs(A[i * n + k]); p = i; }
    piv[k] = p;
    if (p != k) {
      for (int j = 0; j < n; ++j) std::swap(A[k * n + j], A[p * n + j]);
      dt = -dt;
    }
    double akk = A[k * n + k];
    dt *= akk;
    if (akk == 0) { ok = false; continue; }
    for (int i = k + 1; i < n; ++i) {
      double lik = A[i * n + k] / akk;
      A[i * n + k] = lik;
      for (int j = k + 1; j < n; ++j) A[i * n + j] -= lik * A[k * n + j];
    }
  }
  *det = dt;
  return ok;
}
void lu_solve(const double* A, int n, const int* piv, double* b) {
  for (int k = 0; k < n; ++k) std::swap(b[k], b[piv[k]]);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < i; ++j) b[i] -= A[i * n + j] * b[j];
  for (int i = n - 1; i >= 0; --i) {
    for (int j = i + 1; j < n; ++j) b[i] -= A[i * n + j] * b[j];
    b[i] /= A[i * n + i];
  }
}

// ----------------------------------------------------------------------------------------------
// Fantasy surrogate state: rbs.jl:320-333 (X, L, y, cs, observed, fantasies_observed)
// ----------------------------------------------------------------------------------------------
struct FS {
  int d = 0, N = 0, h = 0, ld = 0, nf = 0;
  std::vector<double> X;                 // d x ld column-major
  std::vector<double> L;                 // ld x ld row-major lower
  std::vector<double> y;                 // ld
  std::vector<std::vector<double>> cs;   // cs[0] = base coefficients, cs[k] after k fantasies (rbs.jl:326,427)
};

struct Ctx {
  const orc_problem* p;
  Kern kern;
  double k0, d2k0;  // psi(0), psi''(0)
};

// Evaluation of the fantasy surrogate at x: rbs.jl:482-581.
struct SX {
  int n = 0, d = 0, fi = 0;
  const double* c = nullptr;
  std::vector<double> x, kx, dkx /* d x n col-major */, bj, aj;
  std::vector<double> w, Dw /* column r contiguous: Dw[r*n + j] */;
  double mu = 0, sigma = 0, fstar = 0;
  std::vector<double> dmu, dsig, dal, Hmu, Hsig, Hal_ref, Hal_true, d2a_dxdth;
  GPart g;
  bool neg_var = false;
};

// level 0: value (mu, sigma, alpha); 1: + gradients; 2: + Hessians (+ Dw)
void eval_fs(const Ctx& cx, const FS& fs, const double* x, const double* theta, int fantasy_index, int level, SX& s) {
  const orc_problem* p = cx.p;
  const int d = fs.d, n = fs.N + fantasy_index + 1;  // rbs.jl:499 slice = 1:observed+fantasy_index+1
  const double* c = fs.cs[fantasy_index + 1].data();  // rbs.jl:505 cs[fantasy_index + TOTAL_OFFSET] (1-based)
  s.n = n; s.d = d; s.fi = fantasy_index; s.c = c;
  s.x.assign(x, x + d);
  s.kx.resize(n); s.dkx.assign((size_t)d * n, 0.0); s.bj.resize(n); s.aj.resize(n);
  std::vector<double> r(d);
  // rbf.jl:180-208  eval_KxX, eval_gradKxX ; rbf.jl:141-150 coefficients of eval_Hk
  for (int j = 0; j < n; ++j) {
    double rho2 = 0;
    for (int a = 0; a < d; ++a) { r[a] = x[a] - fs.X[(size_t)j * d + a]; rho2 += r[a] * r[a]; }
    double rho = std::sqrt(rho2), ps, dps, d2ps;
    kern_eval(cx.kern, rho, ps, dps, d2ps);
    s.kx[j] = ps;
    if (rho > 0) {
      double b = dps / rho;
      s.bj[j] = b;
      s.aj[j] = (d2ps - b) / rho2;
      for (int a = 0; a < d; ++a) s.dkx[(size_t)j * d + a] = dps * r[a] / rho;  // rbf.jl:202
    } else {
      s.bj[j] = d2ps;  // rbf.jl:149: Hk(0) = psi''(0) I
      s.aj[j] = 0;
    }
  }
  s.mu = dot(s.kx.data(), c, n);  // rbs.jl:513
  // f* = minimum(get_observations(sx)) over the active slice (rbs.jl:506,550; decision_rules.jl:90)
  double fstar = fs.y[0];
  for (int j = 1; j < n; ++j) fstar = std::min(fstar, fs.y[j]);
  s.fstar = fstar;

  const bool factored = (p->flags & ORC_FLAG_FACTORED) != 0;
  const int nr = (level >= 2) ? d + 1 : 1;
  // right-hand sides [kx, dkx'] (rbs.jl:525-526)
  std::vector<double> V((size_t)nr * n);
  for (int j = 0; j < n; ++j) V[j] = s.kx[j];
  if (nr > 1)
    for (int a = 0; a < d; ++a)
      for (int j = 0; j < n; ++j) V[(size_t)(a + 1) * n + j] = s.dkx[(size_t)j * d + a];
  fwd_solve(fs.L.data(), fs.ld, n, V.data(), n, nr);
  std::vector<double> Vf;
  if (factored) Vf = V;
  bwd_solve(fs.L.data(), fs.ld, n, V.data(), n, nr);
  s.w.assign(V.begin(), V.begin() + n);  // rbs.jl:525
  double var;
  if (factored) var = cx.k0 - dot(Vf.data(), Vf.data(), n);
  else var = cx.k0 - dot(s.kx.data(), s.w.data(), n);  // rbs.jl:528 (no sigma_n2, no clamp)
  s.neg_var = !(var >= 0);
  s.sigma = std::sqrt(var);
  s.g = rule_eval(p->rule_id, p->sigma_tol, s.mu, s.sigma, theta[0], fstar);
  if (level < 1) return;

  s.dmu.assign(d, 0.0); s.dsig.assign(d, 0.0); s.dal.assign(d, 0.0);
  for (int j = 0; j < n; ++j)
    for (int a = 0; a < d; ++a) {
      s.dmu[a] += s.dkx[(size_t)j * d + a] * c[j];      // rbs.jl:514
      s.dsig[a] += s.dkx[(size_t)j * d + a] * s.w[j];   // rbs.jl:529
    }
  for (int a = 0; a < d; ++a) s.dsig[a] = -s.dsig[a] / s.sigma;
  for (int a = 0; a < d; ++a) s.dal[a] = s.g.g_mu * s.dmu[a] + s.g.g_sig * s.dsig[a];  // rbs.jl:567
  s.d2a_dxdth.assign(d, 0.0);
  for (int a = 0; a < d; ++a) s.d2a_dxdth[a] = s.dmu[a] * s.g.g_muth + s.dsig[a] * s.g.g_sigth;  // rbs.jl:575-577
  if (level < 2) return;

  s.Dw.assign(V.begin() + n, V.end());  // rbs.jl:526  Dw = L'\(L\dkx')
  if (factored) {
    // the CUDA kernel's formulation of the mean as well: mu = kx.c = (L^-1 kx).(L^-1 y), grad mu = (L^-1 dkx').(L^-1 y); used by the
    // 1e-10 step-level comparison (algebraically rbs.jl:513-514, rounded differently by kappa * eps)
    std::vector<double> u(fs.y.begin(), fs.y.begin() + n);
    fwd_solve(fs.L.data(), fs.ld, n, u.data(), n, 1);
    s.mu = dot(Vf.data(), u.data(), n);
    for (int a = 0; a < d; ++a) s.dmu[a] = dot(Vf.data() + (size_t)(a + 1) * n, u.data(), n);
    s.g = rule_eval(p->rule_id, p->sigma_tol, s.mu, s.sigma, theta[0], fstar);
    for (int a = 0; a < d; ++a) s.dal[a] = s.g.g_mu * s.dmu[a] + s.g.g_sig * s.dsig[a];
    for (int a = 0; a < d; ++a) s.d2a_dxdth[a] = s.dmu[a] * s.g.g_muth + s.dsig[a] * s.g.g_sigth;
  }
  s.Hmu.assign((size_t)d * d, 0.0); s.Hsig.assign((size_t)d * d, 0.0);
  std::vector<double> Hw((size_t)d * d, 0.0);
  // rbs.jl:516-523, 542-545 with eval_Hk rbf.jl:141-150
  for (int j = 0; j < n; ++j) {
    for (int a = 0; a < d; ++a) r[a] = x[a] - fs.X[(size_t)j * d + a];
    double cj = c[j], wj = s.w[j], aj = s.aj[j], bj = s.bj[j];
    for (int a = 0; a < d; ++a) {
      for (int b = 0; b < d; ++b) {
        double hk = aj * r[a] * r[b] + (a == b ? bj : 0.0);
        s.Hmu[a * d + b] += cj * hk;
        Hw[a * d + b] += wj * hk;
      }
    }
  }
  // rbs.jl:541: H = -dsig dsig' - dkx * Dw ; then -= sum w_j Hk ; /= sigma
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      double kd = 0;
      if (factored) kd = dot(Vf.data() + (size_t)(a + 1) * n, Vf.data() + (size_t)(b + 1) * n, n);
      else
        for (int j = 0; j < n; ++j) kd += s.dkx[(size_t)j * d + a] * s.Dw[(size_t)b * n + j];
      s.Hsig[a * d + b] = (-s.dsig[a] * s.dsig[b] - kd - Hw[a * d + b]) / s.sigma;
    }
  s.Hal_ref.assign((size_t)d * d, 0.0); s.Hal_true.assign((size_t)d * d, 0.0);
  for (int a = 0; a < d; ++a)
    for (int b = 0; b < d; ++b) {
      // rbs.jl:568 (Q1: no mixed mu-sigma term)
      double href = s.g.g_mumu * s.dmu[a] * s.dmu[b] + s.g.g_mu * s.Hmu[a * d + b] +
                    s.g.g_sigsig * s.dsig[a] * s.dsig[b] + s.g.g_sig * s.Hsig[a * d + b];
      s.Hal_ref[a * d + b] = href;
      s.Hal_true[a * d + b] = href + s.g.g_musig * (s.dmu[a] * s.dsig[b] + s.dsig[a] * s.dmu[b]);
    }
}

// rbs.jl:431-441 condition!(fs, x, y): insert, K row (rbs.jl:389-401), Cholesky row (rbs.jl:403-420),
// full coefficient re-solve pushed onto cs (rbs.jl:422-429).
int condition_fs(const Ctx& cx, FS& fs, const double* x, double yv) {
  const int d = fs.d, n = fs.N + fs.nf + 1, ld = fs.ld;  // n = new total
  for (int a = 0; a < d; ++a) fs.X[(size_t)(n - 1) * d + a] = x[a];
  fs.y[n - 1] = yv;
  fs.nf += 1;
  double* Ln = fs.L.data() + (size_t)(n - 1) * ld;
  for (int j = 0; j < n - 1; ++j) {
    double rho2 = 0;
    for (int a = 0; a < d; ++a) { double t = x[a] - fs.X[(size_t)j * d + a]; rho2 += t * t; }
    double ps, dps, d2ps;
    kern_eval(cx.kern, std::sqrt(rho2), ps, dps, d2ps);
    Ln[j] = ps;  // K[n, 1:n-1]
  }
  // L21 = B / L'  (forward substitution), L22 = chol(C - L21 L21')
  fwd_solve(fs.L.data(), ld, n - 1, Ln, n - 1, 1);
  double s = (cx.k0 + cx.p->sigma_n2) - dot(Ln, Ln, n - 1);
  int st = ORC_OK;
  if (!(s > 0)) st = ORC_NOT_PD_ROW;
  Ln[n - 1] = std::sqrt(s);
  std::vector<double> cnew(fs.y.begin(), fs.y.begin() + n);
  fwd_solve(fs.L.data(), ld, n, cnew.data(), n, 1);
  bwd_solve(fs.L.data(), ld, n, cnew.data(), n, 1);
  fs.cs.push_back(std::move(cnew));
  return st;
}

void reset_fs(FS& fs) {  // rbs.jl:476-480
  fs.nf = 0;
  fs.cs.resize(1);
}

// GaussHermiteObservable functor (observables.jl:54-64): y = mu + sqrt(2) sigma node, grad y = grad mu + sqrt(2) grad sigma node
int gh_draw(const Ctx& cx, const FS& fs, const double* x, const double* theta, int fantasy_index, double node, double* out) {
  SX s;
  eval_fs(cx, fs, x, theta, fantasy_index, 1, s);
  const double sqrt2 = 1.4142135623730951;
  out[0] = s.mu + sqrt2 * s.sigma * node;
  for (int a = 0; a < fs.d; ++a) out[1 + a] = s.dmu[a] + sqrt2 * s.dsig[a] * node;
  return s.neg_var ? ORC_NEG_VARIANCE : ORC_OK;
}

// gp_draw with gradient: rbs.jl:588-611 using sx.dmu (rbs.jl:515) and sx.dsigma (rbs.jl:530-539).
// out[0] = y, out[1..d] = grad y.
int gp_draw(const Ctx& cx, const FS& fs, const double* x, const double* theta, int fantasy_index, const double* z, double* out) {
  SX s;
  eval_fs(cx, fs, x, theta, fantasy_index, 1, s);
  const int d = fs.d, n = s.n, q = d + 1;
  int st = ORC_OK;
  if (s.neg_var) st = ORC_NEG_VARIANCE;
  // kxX = [kx'; dkx] ((d+1) x n) ; Sigma = Dk(0) - kxX * (L'\(L\kxX'))   (rbs.jl:531-536)
  std::vector<double> W((size_t)q * n);
  for (int j = 0; j < n; ++j) W[j] = s.kx[j];
  for (int a = 0; a < d; ++a)
    for (int j = 0; j < n; ++j) W[(size_t)(a + 1) * n + j] = s.dkx[(size_t)j * d + a];
  std::vector<double> A = W;
  const bool factored = (cx.p->flags & ORC_FLAG_FACTORED) != 0;
  fwd_solve(fs.L.data(), fs.ld, n, W.data(), n, q);
  if (factored) A = W; else bwd_solve(fs.L.data(), fs.ld, n, W.data(), n, q);
  std::vector<double> Sg((size_t)q * q), Lc((size_t)q * q);
  for (int i = 0; i < q; ++i)
    for (int j = 0; j < q; ++j) {
      // eval_Dk(kernel, zeros(d)) rbf.jl:152-159 = [psi(0) 0; 0 -psi''(0) I]
      double dk = (i == j) ? (i == 0 ? cx.k0 : -cx.d2k0) : 0.0;
      Sg[i * q + j] = dk - dot(A.data() + (size_t)i * n, W.data() + (size_t)j * n, n);
    }
  if (!chol_from_upper(Sg.data(), q, Lc.data())) { st = st ? st : ORC_NOT_PD_JOINT; }
  for (int i = 0; i < q; ++i) {
    double v = (i == 0) ? s.mu : s.dmu[i - 1];
    for (int j = 0; j <= i; ++j) v += Lc[i * q + j] * z[j];
    out[i] = v;
  }
  return st;
}

// ----------------------------------------------------------------------------------------------
// Inner solve. Semantics of rbf_optim.jl:68-101 (S' starts, discard NaN candidates, first minimum
// of -alpha wins); the per-start optimiser replaces Optim.IPNewton (see file header).
// ----------------------------------------------------------------------------------------------
struct StartResult {
  std::vector<double> x;
  double f;
  int status, iters, evals;
};

// ---- exact trust-region subproblem (solve_tr, optim.jl:9-51) without an eigendecomposition ------------------------
// Householder tridiagonalisation H = Q T Q' (Golub & Van Loan 8.3.1), then every quantity of the secular equation
// is an O(n) recurrence on T. The CUDA kernel runs the same steps with one warp (lane = row / candidate shift).

// A: n x n symmetric row-major, overwritten: on exit the diagonal a[] / off-diagonal e[] of T are in ta[0..n), te[0..n-1),
// the Householder vectors v_k (k = 0..n-3) in column k below the sub-diagonal (rows k+1..n-1) and beta[k].
void householder_tridiag(double* A, int n, double* ta, double* te, double* beta, double* wk /* 2n */) {
  double* v = wk; double* pw = wk + n;
  for (int k = 0; k + 2 < n; ++k) {
    double xn2 = 0;
    for (int i = k + 1; i < n; ++i) xn2 += A[i * n + k] * A[i * n + k];
    const double x0 = A[(k + 1) * n + k], alpha = (x0 > 0 ? -1.0 : 1.0) * std::sqrt(xn2);
    const double vtv = xn2 - 2.0 * alpha * x0 + alpha * alpha;  // |x - alpha e1|^2
    if (!(vtv > 0) || !(xn2 - x0 * x0 > 0)) { beta[k] = 0; continue; }  // column already reduced
    const double bk = 2.0 / vtv;
    for (int i = 0; i < n; ++i) v[i] = (i <= k) ? 0.0 : A[i * n + k];
    v[k + 1] = x0 - alpha;
    double pv = 0;
    for (int i = k + 1; i < n; ++i) { double t = 0; for (int j = k + 1; j < n; ++j) t += A[i * n + j] * v[j]; pw[i] = bk * t; pv += pw[i] * v[i]; }
    const double kk = 0.5 * bk * pv;
    for (int i = k + 1; i < n; ++i) pw[i] -= kk * v[i];
    for (int i = k + 1; i < n; ++i)
      for (int j = k + 1; j < n; ++j) A[i * n + j] -= v[i] * pw[j] + pw[i] * v[j];
    A[(k + 1) * n + k] = alpha; A[k * n + (k + 1)] = alpha;
    for (int i = k + 2; i < n; ++i) { A[i * n + k] = v[i]; A[k * n + i] = 0.0; }
    // v[k+1] is kept in beta's companion slot
    beta[k] = bk; beta[n + k] = v[k + 1];
  }
  for (int i = 0; i < n; ++i) ta[i] = A[i * n + i];
  for (int i = 0; i + 1 < n; ++i) te[i] = A[(i + 1) * n + i];
}
// y <- H_k y for k ascending (Q' y, transpose = true) or descending (Q y): H_k = I - beta_k v_k v_k'
void apply_reflectors(const double* A, const double* beta, int n, double* y, bool transpose) {
  for (int kk = 0; kk + 2 < n; ++kk) {
    const int k = transpose ? kk : n - 3 - kk;
    if (beta[k] == 0) continue;
    double dot = beta[n + k] * y[k + 1];
    for (int i = k + 2; i < n; ++i) dot += A[i * n + k] * y[i];
    dot *= beta[k];
    y[k + 1] -= dot * beta[n + k];
    for (int i = k + 2; i < n; ++i) y[i] -= dot * A[i * n + k];
  }
}
// Forward recurrences of (T + lam I) = L D L', z = L^-1 b and their lam-derivatives: returns false if a pivot is <= 0 (not
// positive definite), else |h(lam)|^2 = b'(T + lam I)^-2 b = -d/dlam [sum z_i^2 / d_i].
bool tri_norm2(const double* ta, const double* te, const double* b, int n, double lam, double* hn2) {
  double dprev = ta[0] + lam, dd = 1.0, z = b[0], zd = 0.0;
  if (!(dprev > 0)) return false;
  double r = 1.0 / dprev;
  double acc = (z * z * dd) * r * r;  // -(d/dlam)(z^2/d) = (z^2 d' - 2 z z' d) / d^2
  for (int i = 1; i < n; ++i) {
    const double e = te[i - 1], er = e * r;
    const double dn = ta[i] + lam - e * er, ddn = 1.0 + er * er * dd;
    const double zn = b[i] - er * z, zdn = -er * zd + er * r * dd * z;
    if (!(dn > 0)) return false;
    r = 1.0 / dn;
    acc += (zn * zn * ddn - 2.0 * zn * zdn * dn) * r * r;
    dprev = dn; dd = ddn; z = zn; zd = zdn;
  }
  *hn2 = acc;
  return true;
}
// h = -(T + lam I)^-1 b (requires positive pivots)
void tri_solve_neg(const double* ta, const double* te, const double* b, int n, double lam, double* h, double* wk /* 2n */) {
  double* r = wk; double* z = wk + n;
  r[0] = 1.0 / (ta[0] + lam); z[0] = b[0];
  for (int i = 1; i < n; ++i) { const double er = te[i - 1] * r[i - 1]; r[i] = 1.0 / (ta[i] + lam - te[i - 1] * er); z[i] = b[i] - er * z[i - 1]; }
  h[n - 1] = z[n - 1] * r[n - 1];
  for (int i = n - 2; i >= 0; --i) h[i] = (z[i] - te[i] * h[i + 1]) * r[i];
  for (int i = 0; i < n; ++i) h[i] = -h[i];
}

// Exact trust-region step: minimise g'p + p'Hp/2 subject to |p|_2 <= Delta (solve_tr, optim.jl:9-51). The admissible
// shifts S = {lam >= 0 : T + lam I positive definite and |h(lam)| <= Delta} form a half-line [lam*, inf): lam* = 0 is the
// interior Newton step (optim.jl:13-21), otherwise the boundary solution, and in the hard case lam* = -lambda_min with
// |h(lam*)| < Delta, completed along the lowest eigenvector (optim.jl:39-46). lam* is located by TR_ROUNDS rounds of
// TR_CAND-way multisection (each lane of the CUDA warp tests two candidates), i.e. to 1/(63 * 64^2) = 4e-6 of the initial
// bracket: a trust-region step does not need |p| = Delta to more than that. Hff: n x n row-major (overwritten).
// Returns true when the constraint is active ("hit_constraint").
constexpr int TR_ROUNDS = 3, TR_CAND = 64;
bool tr_step(double* Hff, const double* gf, int n, double Delta, double* pout, double* work /* >= 8 n */) {
  double* ta = work; double* te = ta + n; double* beta = te + n; double* gt = beta + 2 * n; double* wk = gt + n;  // wk: 2n
  if (n == 1) {
    const double hh = Hff[0], g0 = gf[0];
    if (hh > 0 && std::fabs(g0 / hh) <= Delta) { pout[0] = -g0 / hh; return false; }
    pout[0] = (g0 > 0 ? -Delta : Delta);
    return true;
  }
  householder_tridiag(Hff, n, ta, te, beta, wk);
  double gn2 = 0;
  for (int i = 0; i < n; ++i) { gt[i] = gf[i]; gn2 += gf[i] * gf[i]; }
  apply_reflectors(Hff, beta, n, gt, true);
  double gl = ta[0] - std::fabs(te[0]);  // Gershgorin lower bound of lambda_min(T)
  for (int i = 1; i < n; ++i) gl = std::min(gl, ta[i] - std::fabs(te[i - 1]) - (i + 1 < n ? std::fabs(te[i]) : 0.0));
  const double D2 = Delta * Delta;
  // 0: admissible; 1: positive definite but |h| > Delta; 2: not positive definite
  auto probe = [&](double lam) { double hn2; if (!tri_norm2(ta, te, gt, n, lam, &hn2)) return 2; return hn2 <= D2 ? 0 : 1; };
  double lo = 0.0, hi = std::max(0.0, -gl) + std::sqrt(gn2) / Delta;
  bool hit = true, lo_notpd = false;
  const int p0 = probe(0.0);
  if (p0 == 0) { hi = 0.0; hit = false; }
  else {
    lo_notpd = p0 == 2;
    // TR_CAND candidates per round (two per lane of the CUDA warp); the last one is hi, known to be admissible
    for (int round = 0; round < TR_ROUNDS; ++round) {
      const double base = lo, wd = hi - lo;
      int cfirst = TR_CAND - 1;
      bool below_notpd = lo_notpd;  // state of the candidate just below the first admissible one
      for (int c = (round == 0 ? 1 : 0); c < TR_CAND - 1; ++c) {
        const int pr = probe(round == 0 ? wd * (double)c / (TR_CAND - 1) : base + wd * (double)(c + 1) / TR_CAND);
        if (pr == 0) { cfirst = c; break; }
        below_notpd = pr == 2;
      }
      lo_notpd = below_notpd;
      if (round == 0) {
        hi = (cfirst == TR_CAND - 1) ? hi : wd * (double)cfirst / (TR_CAND - 1);
        lo = wd * (double)(cfirst - 1) / (TR_CAND - 1);
      } else {
        hi = (cfirst == TR_CAND - 1) ? hi : base + wd * (double)(cfirst + 1) / TR_CAND;
        lo = (cfirst == 0) ? base : base + wd * (double)cfirst / TR_CAND;
      }
    }
  }
  tri_solve_neg(ta, te, gt, n, hi, pout, wk);
  if (hit && lo_notpd) {
    double hn2 = 0;
    for (int i = 0; i < n; ++i) hn2 += pout[i] * pout[i];
    if (hn2 < D2) {
      // the admissible set ends where T + lam I stops being positive definite and the step there is still short: hard case
      // (to the resolution of the search); complete along the lowest eigenvector of T (two steps of inverse iteration at
      // the barely definite shift) -- this only lowers the model further (optim.jl:39-46)
      std::vector<double> z(n, 1.0), z2(n);
      for (int itn = 0; itn < 2; ++itn) {
        tri_solve_neg(ta, te, z.data(), n, hi, z2.data(), wk);
        double zn = 0;
        for (int i = 0; i < n; ++i) zn += z2[i] * z2[i];
        zn = 1.0 / std::sqrt(zn);
        for (int i = 0; i < n; ++i) z[i] = z2[i] * zn;
      }
      double hz = 0;
      for (int i = 0; i < n; ++i) hz += pout[i] * z[i];
      const double tau = -hz + std::sqrt(std::max(hz * hz + (D2 - hn2), 0.0));  // |h + tau z| = Delta, |z| = 1
      for (int i = 0; i < n; ++i) pout[i] += tau * z[i];
    }
  }
  apply_reflectors(Hff, beta, n, pout, false);
  return hit;
}

// One start of the inner solve: trust-region Newton (tr_newton, optim.jl:68-114, with the exact subproblem solver above) on
// the merit f = -log(alpha) (EI, POI while alpha > 0; the argmax is the same and the tails of EI become near-quadratic) or
// f = -alpha (LCB), projected onto the box with a gradient-sign active set. Specified in DESIGN.md section 4; the CUDA kernel
// (slot_logic_warp) implements the same algorithm independently.
void solve_start(const Ctx& cx, const FS& fs, const double* theta, int fantasy_index, const double* start, StartResult& res) {
  const orc_problem* p = cx.p;
  const orc_solver_opts& o = p->solver;
  const int d = fs.d;
  std::vector<double> x(d), g(d), H((size_t)d * d), xt(d), sv(d), A((size_t)d * d), pv(d), gf(d), work(8 * (size_t)d + 8);
  std::vector<int> fr(d);
  SX s, st;
  for (int a = 0; a < d; ++a) x[a] = std::min(std::max(start[a], p->lbs[a]), p->ubs[a]);
  eval_fs(cx, fs, x.data(), theta, fantasy_index, 2, s);
  res.evals = 1; res.iters = 0;
  double alpha = s.g.g;
  const bool logm = p->rule_id != ORC_RULE_LCB && alpha > 0;  // fixed per start
  auto finite_eval = [&](const SX& sx) {
    if (!std::isfinite(sx.g.g)) return false;
    for (int a = 0; a < d; ++a) if (!std::isfinite(sx.dal[a])) return false;
    for (int i = 0; i < d * d; ++i) if (!std::isfinite(sx.Hal_true[i])) return false;
    return true;
  };
  double f = 0;
  auto load = [&](const SX& sx) {
    alpha = sx.g.g;
    if (logm) {
      const double ia = 1.0 / alpha;
      f = -std::log(alpha);
      for (int a = 0; a < d; ++a) g[a] = -sx.dal[a] * ia;
      for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b) H[a * d + b] = -sx.Hal_true[a * d + b] * ia + (sx.dal[a] * ia) * (sx.dal[b] * ia);
    } else {
      f = -alpha;
      for (int a = 0; a < d; ++a) g[a] = -sx.dal[a];
      for (int i = 0; i < d * d; ++i) H[i] = -sx.Hal_true[i];
    }
  };
  res.status = ORC_SOLVE_MAXIT;
  if (!finite_eval(s)) { res.status = ORC_SOLVE_NAN; res.x = x; res.f = std::numeric_limits<double>::quiet_NaN(); return; }
  load(s);
  double wmax = 0, dmax2 = 0;
  for (int a = 0; a < d; ++a) { const double wd = p->ubs[a] - p->lbs[a]; wmax = std::max(wmax, wd); dmax2 += wd * wd; }
  const double Dmax = std::sqrt(dmax2);
  double Delta = std::min(o.delta0_box * wmax, o.delta0_ell * cx.kern.th[0]);
  int tries = 0;
  for (;;) {
    int nfree = 0;
    double pg = 0;  // projected gradient of alpha itself (the stopping test does not depend on the merit)
    for (int a = 0; a < d; ++a) {
      const bool act = (x[a] <= p->lbs[a] && g[a] > 0) || (x[a] >= p->ubs[a] && g[a] < 0);
      if (!act) { fr[nfree++] = a; pg = std::max(pg, std::fabs(g[a])); }
    }
    if (logm) pg *= alpha;
    if (pg <= o.gtol * std::max(1.0, std::fabs(alpha))) { res.status = ORC_SOLVE_CONVERGED; break; }
    for (int i = 0; i < nfree; ++i) {
      gf[i] = g[fr[i]];
      for (int j = 0; j < nfree; ++j) A[i * nfree + j] = H[fr[i] * d + fr[j]];
    }
    const bool hit = tr_step(A.data(), gf.data(), nfree, Delta, pv.data(), work.data());
    for (int a = 0; a < d; ++a) xt[a] = x[a];
    for (int i = 0; i < nfree; ++i) { const int a = fr[i]; xt[a] = std::min(std::max(x[a] + pv[i], p->lbs[a]), p->ubs[a]); }
    double smax = 0, xmax = 0, sn2 = 0;
    for (int a = 0; a < d; ++a) { sv[a] = xt[a] - x[a]; smax = std::max(smax, std::fabs(sv[a])); xmax = std::max(xmax, std::fabs(x[a])); sn2 += sv[a] * sv[a]; }
    const double sn = std::sqrt(sn2);
    if (smax <= o.xtol * std::max(1.0, xmax)) { res.status = ORC_SOLVE_STEP_TINY; break; }
    double gs = 0, sHs = 0;
    for (int a = 0; a < d; ++a) {
      gs += g[a] * sv[a];
      double t = 0;
      for (int b = 0; b < d; ++b) t += H[a * d + b] * sv[b];
      sHs += sv[a] * t;
    }
    const double pred = -(gs + 0.5 * sHs);  // mu_diff of optim.jl:88, for the projected step
    if (!(pred > 0)) {
      Delta = 0.25 * std::min(Delta, sn);
      if (++tries >= o.maxtry) { res.status = ORC_SOLVE_STALLED; break; }
      continue;
    }
    const bool pred_tiny = pred <= o.pred_tol * std::max(1.0, std::fabs(f));
    if (!hit && (pred_tiny || smax <= o.stol * std::max(1.0, xmax))) {
      // final interior Newton step: taken without another evaluation (its error is O(|s|^2)); alpha follows the model
      x = xt; alpha = logm ? alpha * std::exp(pred) : alpha + pred; res.status = ORC_SOLVE_FINAL_STEP; break;
    }
    if (pred_tiny) { res.status = ORC_SOLVE_PRED_TINY; break; }
    eval_fs(cx, fs, xt.data(), theta, fantasy_index, 2, st);
    res.evals++;
    const bool fin = finite_eval(st) && (!logm || st.g.g > 0);
    const double ft = fin ? (logm ? -std::log(st.g.g) : -st.g.g) : 0.0;
    const double rho = fin ? (f - ft) / pred : -1.0;
    if (fin && rho >= o.eta) {  // optim.jl:99
      x = xt; load(st);
      if (rho > 0.75 && hit && sn >= 0.8 * Delta) Delta = std::min(2.0 * Delta, Dmax);  // optim.jl:95-96 (projection may have shortened the step)
      else if (rho < 0.25) Delta = 0.25 * sn;                                               // optim.jl:93-94
      tries = 0;
      if (++res.iters >= o.maxit) { res.status = ORC_SOLVE_MAXIT; break; }
    } else {
      Delta = 0.25 * std::min(Delta, sn);
      if (++tries >= o.maxtry) { res.status = ORC_SOLVE_STALLED; break; }
    }
  }
  res.x = x; res.f = -alpha;
}

int multistart(const Ctx& cx, const FS& fs, const double* theta, int fantasy_index, double* xbest, double* fbest, int* evals,
               int* st_status, int* st_iters, double* st_x, double* st_f) {
  const orc_problem* p = cx.p;
  const int d = fs.d;
  int best = -1, ne = 0;
  double fb = 0;
  StartResult r;
  for (int s = 0; s < p->S; ++s) {
    solve_start(cx, fs, theta, fantasy_index, p->starts + (size_t)s * d, r);
    ne += r.evals;
    if (st_status) st_status[s] = r.status;
    if (st_iters) st_iters[s] = r.iters;
    if (st_f) st_f[s] = r.f;
    if (st_x) for (int a = 0; a < d; ++a) st_x[(size_t)s * d + a] = r.x[a];
    bool nan = !std::isfinite(r.f);
    for (int a = 0; a < d; ++a) if (std::isnan(r.x[a])) nan = true;
    if (nan) continue;                       // rbf_optim.jl:96
    if (best < 0 || r.f < fb) {              // rbf_optim.jl:97 findmin: first minimum wins
      best = s; fb = r.f;
      for (int a = 0; a < d; ++a) xbest[a] = r.x[a];
    }
  }
  if (evals) *evals = ne;
  if (fbest) *fbest = fb;
  return best < 0 ? ORC_ALL_STARTS_NAN : ORC_OK;
}

// ----------------------------------------------------------------------------------------------
// Perturbation surrogates: rbs.jl:633-764 with rbf.jl:210-262
// ----------------------------------------------------------------------------------------------
struct Perturb {
  std::vector<double> dal;  // delta grad alpha
};

// sx: evaluation at x_i (n = N+i). Moves column pcol (0-based) of the active set by dx.
// spatial=true -> rbs.jl:652-694 ; spatial=false -> rbs.jl:711-760 (Q7: no g_sigma * delta grad sigma term)
void eval_perturbation(const Ctx& cx, const FS& fs, const SX& sx, const double* theta, int pcol, const double* dx, bool spatial, std::vector<double>& out) {
  const orc_problem* p = cx.p;
  const int d = fs.d, n = sx.n;
  const double* c = sx.c;  // cs[max_fantasized_step + TOTAL_OFFSET] == coefficients sx was built with (rbs.jl:675)
  const bool fast = (p->flags & ORC_FLAG_FAST_PERTURB) != 0;
  std::vector<double> dKc(n, 0.0), dKw(n, 0.0), r(d);
  const double* Xp = fs.X.data() + (size_t)pcol * d;
  if (!fast) {
    // rbf.jl:210-228: dense dK with dK_ij = grad_k(X_i - X_j) . (dX_i - dX_j), only column pcol of dX non-zero
    std::vector<double> dK((size_t)n * n, 0.0);
    for (int j = 0; j < n; ++j)
      for (int i = j + 1; i < n; ++i) {
        double rho2 = 0;
        for (int a = 0; a < d; ++a) { r[a] = fs.X[(size_t)i * d + a] - fs.X[(size_t)j * d + a]; rho2 += r[a] * r[a]; }
        double rho = std::sqrt(rho2), v = 0;
        if (rho != 0) {
          double ps, dps, d2ps;
          kern_eval(cx.kern, rho, ps, dps, d2ps);
          double t = 0;
          for (int a = 0; a < d; ++a) {
            double dXi = (i == pcol) ? dx[a] : 0.0, dXj = (j == pcol) ? dx[a] : 0.0;
            t += (dps * r[a] / rho) * (dXi - dXj);
          }
          v = t;
        }
        dK[(size_t)i * n + j] = v;
        dK[(size_t)j * n + i] = v;
      }
    for (int i = 0; i < n; ++i) {
      dKc[i] = dot(dK.data() + (size_t)i * n, c, n);
      dKw[i] = dot(dK.data() + (size_t)i * n, sx.w.data(), n);
    }
  } else {
    // rank-2 structure: u_a = grad_k(X_a - X_p).(-dx) for a != p
    double uc = 0, uw = 0;
    for (int a_ = 0; a_ < n; ++a_) {
      if (a_ == pcol) continue;
      double rho2 = 0;
      for (int a = 0; a < d; ++a) { r[a] = fs.X[(size_t)a_ * d + a] - Xp[a]; rho2 += r[a] * r[a]; }
      double rho = std::sqrt(rho2), u = 0;
      if (rho != 0) {
        double ps, dps, d2ps;
        kern_eval(cx.kern, rho, ps, dps, d2ps);
        for (int a = 0; a < d; ++a) u += (dps * r[a] / rho) * (-dx[a]);
      }
      dKc[a_] = u * c[pcol];
      dKw[a_] = u * sx.w[pcol];
      uc += u * c[a_];
      uw += u * sx.w[a_];
    }
    dKc[pcol] = uc;
    dKw[pcol] = uw;
  }
  // dc = -(L'\(L\(dK c)))  rbs.jl:675
  std::vector<double> dc = dKc;
  fwd_solve(fs.L.data(), fs.ld, n, dc.data(), n, 1);
  bwd_solve(fs.L.data(), fs.ld, n, dc.data(), n, 1);
  for (int j = 0; j < n; ++j) dc[j] = -dc[j];
  // dkx (rbf.jl:230-245) and d(grad kx) (rbf.jl:247-262): only entry/column pcol is non-zero
  double dkx_p = 0;
  std::vector<double> dgkx_p(d, 0.0);
  {
    double rho2 = 0;
    for (int a = 0; a < d; ++a) { r[a] = sx.x[a] - Xp[a]; rho2 += r[a] * r[a]; }
    double rho = std::sqrt(rho2), ps, dps, d2ps;
    kern_eval(cx.kern, rho, ps, dps, d2ps);
    if (rho != 0)
      for (int a = 0; a < d; ++a) dkx_p += (dps * r[a] / rho) * (-dx[a]);
    // Hk(x - X_p) * (-dx)
    for (int a = 0; a < d; ++a) {
      double t = 0;
      for (int b = 0; b < d; ++b) {
        double hk;
        if (rho > 0) {
          double Dpr = dps / rho;
          hk = (d2ps - Dpr) * (r[a] / rho) * (r[b] / rho) + (a == b ? Dpr : 0.0);
        } else hk = (a == b) ? d2ps : 0.0;
        t += hk * (-dx[b]);
      }
      dgkx_p[a] = t;
    }
  }
  // dmu (rbs.jl:680), dgrad mu (rbs.jl:681)
  double dmu = dkx_p * c[pcol] + dot(sx.kx.data(), dc.data(), n);
  std::vector<double> dgmu(d, 0.0), dgsig(d, 0.0);
  for (int a = 0; a < d; ++a) {
    double t = dgkx_p[a] * c[pcol];
    for (int j = 0; j < n; ++j) t += sx.dkx[(size_t)j * d + a] * dc[j];
    dgmu[a] = t;
  }
  // dsigma (rbs.jl:683)
  double dsig = (-2 * dkx_p * sx.w[pcol] + dot(sx.w.data(), dKw.data(), n)) / (2 * sx.sigma);
  // dgrad sigma (rbs.jl:684): (Dw'(dK w) - dgkx w - Dw' dkx - dsigma grad sigma) / sigma
  if (spatial)
    for (int a = 0; a < d; ++a) {
      double t = dot(sx.Dw.data() + (size_t)a * n, dKw.data(), n) - dgkx_p[a] * sx.w[pcol] - sx.Dw[(size_t)a * n + pcol] * dkx_p - dsig * sx.dsig[a];
      dgsig[a] = t / sx.sigma;
    }
  // Q6: partials evaluated AT (dmu, dsigma) (rbs.jl:687-688), f* from sx
  GPart gh = rule_eval(p->rule_id, p->sigma_tol, dmu, dsig, theta[0], sx.fstar);
  out.assign(d, 0.0);
  for (int a = 0; a < d; ++a) {
    if (spatial)  // rbs.jl:690
      out[a] = sx.g.g_mu * dgmu[a] + sx.g.g_sig * dgsig[a] + gh.g_mu * sx.dmu[a] + gh.g_sig * sx.dsig[a];
    else          // rbs.jl:756
      out[a] = sx.g.g_mu * dgmu[a] + gh.g_mu * sx.dmu[a] + gh.g_sig * sx.dsig[a];
  }
}

// ----------------------------------------------------------------------------------------------
// Adjoint gradient of one trajectory: rollout.jl:114-277
// ----------------------------------------------------------------------------------------------
struct Traj {
  std::vector<double> obs;    // observations[k], k = 0..h  (observables.jl:16-19)
  std::vector<double> grads;  // gradients[:, k]
};

// rollout.jl:114-124: sx at x_j (column N+j) with fantasy_index = j-1
void recover_policy_solve(const Ctx& cx, const FS& fs, const double* theta, int solve_index, SX& s) {
  eval_fs(cx, fs, fs.X.data() + (size_t)(fs.N + solve_index) * fs.d, theta, solve_index - 1, 2, s);
}

int trajectory_gradient(const Ctx& cx, const FS& fs, const Traj& tj, const double* theta, const double* dual_dirs_m /* d x h */, double* gx, double* gth, int* tcase, int* tbest) {
  const orc_problem* p = cx.p;
  const int d = fs.d, h = fs.h, nth = p->ntheta;
  int status = ORC_OK;
  for (int a = 0; a < d; ++a) gx[a] = 0;
  for (int a = 0; a < nth; ++a) gth[a] = 0;
  // rollout.jl:77-105: findmin over the fantasy observations (first minimum), t = index - 1
  int t = 0;
  double fb = fs.y[fs.N];
  for (int k = 1; k <= h; ++k)
    if (fs.y[fs.N + k] < fb) { fb = fs.y[fs.N + k]; t = k; }
  *tbest = t;
  if (p->fmini <= fb) { *tcase = 1; return status; }                   // rollout.jl:241-243
  if (t == 0) {                                                        // rollout.jl:249
    *tcase = 2;
    for (int a = 0; a < d; ++a) gx[a] = -tj.grads[a];
    return status;
  }
  *tcase = 3;
  std::vector<std::vector<double>> xbars(t + 1, std::vector<double>(d, 0.0));  // 1-based j
  std::vector<double> ybars(t + 2, 0.0);                                        // ybars[j], j = 1..t+1
  ybars[t + 1] = 1.0;                                                           // rollout.jl:256
  SX sx, sxi;
  std::vector<double> dal, e(d), Hm((size_t)d * d), rhs(d), dri((size_t)d * d);
  std::vector<int> piv(d);
  for (int j = t; j >= 1; --j) {
    // ---- solve_dual_x (rollout.jl:150-191)
    recover_policy_solve(cx, fs, theta, j, sx);
    for (int a = 0; a < d; ++a)
      for (int b = 0; b < d; ++b) Hm[a * d + b] = sx.Hal_ref[a * d + b];
    std::vector<double> Hlu = Hm;
    double det;
    lu_factor(Hlu.data(), d, piv.data(), &det);
    if (det < p->htol) {                                               // rollout.jl:159-161 (Q3)
      std::fill(xbars[j].begin(), xbars[j].end(), 0.0);
    } else {
      for (int a = 0; a < d; ++a) rhs[a] = -tj.grads[(size_t)(j - 1) * d + a] * ybars[j + 1];  // rollout.jl:164-165 (Q4: at = j)
      for (int i = j + 1; i <= t; ++i) {                               // rollout.jl:173-186
        recover_policy_solve(cx, fs, theta, i, sxi);
        for (int k = 0; k < d; ++k) {
          std::fill(e.begin(), e.end(), 0.0); e[k] = 1.0;
          eval_perturbation(cx, fs, sxi, theta, fs.N + j, e.data(), true, dal);
          for (int a = 0; a < d; ++a) dri[a * d + k] = dal[a];        // dri_dxj[:, k]
        }
        for (int k = 0; k < d; ++k) {                                  // x_dual -= dri' * xbars[i]
          double s = 0;
          for (int a = 0; a < d; ++a) s += dri[a * d + k] * xbars[i][a];
          rhs[k] -= s;
        }
      }
      // x_dual = hessian(sx)' \ x_dual   (rollout.jl:188)
      std::vector<double> Ht((size_t)d * d);
      for (int a = 0; a < d; ++a)
        for (int b = 0; b < d; ++b) Ht[a * d + b] = Hm[b * d + a];
      double det2;
      if (!lu_factor(Ht.data(), d, piv.data(), &det2)) status = ORC_SINGULAR_HESSIAN;
      lu_solve(Ht.data(), d, piv.data(), rhs.data());
      xbars[j] = rhs;
    }
    // ---- solve_dual_y (rollout.jl:126-148), solve_index = j-1; dx = rand(dim) supplied by the caller (Q5)
    {
      int sidx = j - 1;
      const double* dxr = dual_dirs_m + (size_t)sidx * d;
      double yd = 0;
      for (int i = sidx + 1; i <= t; ++i) {
        recover_policy_solve(cx, fs, theta, i, sxi);
        eval_perturbation(cx, fs, sxi, theta, fs.N + sidx, dxr, false, dal);
        for (int a = 0; a < d; ++a) yd += dal[a] * xbars[i][a];
      }
      ybars[j] = yd;
    }
  }
  // gather_g (rollout.jl:193-218) and gather_q (rollout.jl:220-231); final assembly rollout.jl:267-276
  recover_policy_solve(cx, fs, theta, 0, sx);  // x_0 under the base GP (fantasy_index = -1)
  std::vector<double> grad_x(d, 0.0), grad_th(nth, 0.0);
  for (int a = 0; a < d; ++a) grad_x[a] = sx.dmu[a] * ybars[1];
  for (int j = 1; j <= t; ++j) {
    recover_policy_solve(cx, fs, theta, j, sxi);
    for (int k = 0; k < d; ++k) {
      std::fill(e.begin(), e.end(), 0.0); e[k] = 1.0;
      eval_perturbation(cx, fs, sxi, theta, fs.N + 0, e.data(), true, dal);
      double s = 0;
      for (int a = 0; a < d; ++a) s += dal[a] * xbars[j][a];          // (g[j+1]' * xbars[j])[k]
      grad_x[k] += s;
    }
    // q[j]' * xbars[j]: q = d2alpha/dx dtheta (d x ntheta); only theta[1] enters the rules
    double s = 0;
    for (int a = 0; a < d; ++a) s += sxi.d2a_dxdth[a] * xbars[j][a];
    grad_th[0] += s;
  }
  for (int a = 0; a < d; ++a) gx[a] = -grad_x[a];
  for (int a = 0; a < nth; ++a) gth[a] = -grad_th[a];
  return status;
}

void setup_ctx(const orc_problem* p, Ctx& cx) {
  cx.p = p;
  cx.kern.id = p->kernel_id;
  for (int i = 0; i < 4; ++i) cx.kern.th[i] = (i < p->nktheta) ? p->ktheta[i] : 0.0;
  double a, b;
  kern_eval(cx.kern, 0.0, cx.k0, a, b);
  cx.d2k0 = b;
}

void init_fs(const orc_problem* p, FS& fs) {
  fs.d = p->d; fs.N = p->N; fs.h = p->h; fs.ld = p->N + p->h + 1; fs.nf = 0;
  fs.X.assign((size_t)fs.d * fs.ld, 0.0);
  fs.L.assign((size_t)fs.ld * fs.ld, 0.0);
  fs.y.assign(fs.ld, 0.0);
  for (int j = 0; j < p->N; ++j)
    for (int a = 0; a < p->d; ++a) fs.X[(size_t)j * fs.d + a] = p->X[(size_t)j * p->ldX + a];
  for (int i = 0; i < p->N; ++i)
    for (int k = 0; k <= i; ++k) fs.L[(size_t)i * fs.ld + k] = p->L[(size_t)k * p->ldL + i];
  for (int j = 0; j < p->N; ++j) fs.y[j] = p->y[j];
  fs.cs.clear();
  fs.cs.emplace_back(p->c, p->c + p->N);
}

}  // namespace

extern "C" {

void orc_default_solver_opts(orc_solver_opts* o) {
  o->maxit = 100;
  o->maxtry = 30;
  o->gtol = 1e-10;
  o->xtol = 1e-15;
  o->pred_tol = 1e-13;
  o->eta = 0.1;
  o->delta0_box = 0.5;
  o->delta0_ell = 1.0;
  o->stol = 1e-5;
}

int orc_rollout(const orc_problem* p, orc_outputs* out) {
  Ctx cx;
  setup_ctx(p, cx);
  const int d = p->d, h = p->h, M = p->M, q = d + 1, nth = p->ntheta;
  const bool want_grad = p->mode == ORC_MODE_VALUE_GRAD && out->grad_x && out->grad_theta;
#ifdef _OPENMP
  int nt = p->nthreads > 0 ? p->nthreads : omp_get_max_threads();
#else
  int nt = 1;
#endif
  (void)nt;
#pragma omp parallel num_threads(nt)
  {
    FS fs;
    init_fs(p, fs);
    Traj tj;
    std::vector<double> z(q), draw(q), xnext(d), zero_dirs((size_t)d * std::max(h, 1), 0.0), gx(d), gth(std::max(nth, 1));
#pragma omp for schedule(dynamic, 1)
    for (int m = 0; m < M; ++m) {
      reset_fs(fs);
      tj.obs.assign(h + 1, 0.0);
      tj.grads.assign((size_t)d * (h + 1), 0.0);
      int status = ORC_OK;
      // rollout! (rollout.jl:39-74)
      for (int step = 0; step <= h; ++step) {
        const double* xloc;
        if (step == 0) xloc = p->x0;  // rollout.jl:46
        else {
          int ne = 0;
          if (p->flags & ORC_FLAG_TEACHER_FORCED) {
            for (int a = 0; a < d; ++a) xnext[a] = p->x_forced[((size_t)m * h + (step - 1)) * d + a];
          } else {
            int* sst = out->start_status ? out->start_status + ((size_t)m * h + (step - 1)) * p->S : nullptr;
            int* sit = out->start_iters ? out->start_iters + ((size_t)m * h + (step - 1)) * p->S : nullptr;
            double fb;
            int ms = multistart(cx, fs, p->theta, step - 1, xnext.data(), &fb, &ne, sst, sit, nullptr, nullptr);  // rollout.jl:58-66
            if (ms != ORC_OK && status == ORC_OK) status = ms;
          }
          if (out->n_evals) out->n_evals[(size_t)m * h + (step - 1)] = ne;
          if (out->alphas) {
            SX sa;
            eval_fs(cx, fs, xnext.data(), p->theta, step - 1, 0, sa);
            out->alphas[(size_t)m * h + (step - 1)] = sa.g.g;
          }
          if (out->t_mu) {  // extended tape: sx = fs(x_step, theta; fantasy_index = step - 1) (rbs.jl:482-581)
            SX sa;
            eval_fs(cx, fs, xnext.data(), p->theta, step - 1, 2, sa);
            const size_t o = (size_t)m * h + (step - 1);
            out->t_mu[o] = sa.mu; out->t_sigma[o] = sa.sigma;
            for (int a = 0; a < d; ++a) { out->t_dmu[o * d + a] = sa.dmu[a]; out->t_dsigma[o * d + a] = sa.dsig[a]; }
            for (int i = 0; i < d * d; ++i) out->t_Halpha[o * d * d + i] = sa.Hal_ref[i];
          }
          xloc = xnext.data();
        }
        // StochasticObservable functor (observables.jl:106-121): z = stdnormal[:, step+1], fantasy_index = step-1
        int ds;
        if (p->flags & ORC_FLAG_GAUSS_HERMITE) {
          ds = gh_draw(cx, fs, xloc, p->theta, step - 1, p->gh_nodes[(size_t)m * (h + 1) + step], draw.data());
        } else {
          for (int k = 0; k < q; ++k) z[k] = p->rn[(size_t)m + (size_t)M * k + (size_t)M * q * step];
          ds = gp_draw(cx, fs, xloc, p->theta, step - 1, z.data(), draw.data());
        }
        if (ds != ORC_OK && status == ORC_OK) status = ds;
        tj.obs[step] = draw[0];
        for (int a = 0; a < d; ++a) tj.grads[(size_t)step * d + a] = draw[1 + a];
        int cs = condition_fs(cx, fs, xloc, draw[0]);  // rollout.jl:49,72
        if (cs != ORC_OK && status == ORC_OK) status = cs;
      }
      // resolve (rollout.jl:108-111; observables.jl:12-14)
      double best = tj.obs[0];
      int kbest = 0;
      for (int k = 1; k <= h; ++k) if (tj.obs[k] < best) { best = tj.obs[k]; kbest = k; }
      out->values[m] = std::max(p->fmini - best, 0.0);
      if (p->flags & ORC_FLAG_GAUSS_HERMITE) {
        // resolve(gho; fmini) (observables.jl:66-72): weight of the best step, 1/sqrt(pi); get_gradient (observables.jl:157,
        // the later definition wins): weights[at] * gradients[:, at]
        out->values[m] = p->gh_weights[(size_t)m * (h + 1) + kbest] * std::max(p->fmini - best, 0.0) / 1.7724538509055159;
        for (int k = 0; k <= h; ++k)
          for (int a = 0; a < d; ++a) tj.grads[(size_t)k * d + a] *= p->gh_weights[(size_t)m * (h + 1) + k];
      }
      int tcase = 0, tbest = 0;
      {
        double fb = tj.obs[0];
        for (int k = 1; k <= h; ++k) if (tj.obs[k] < fb) { fb = tj.obs[k]; tbest = k; }
      }
      if (want_grad) {
        const double* dd = p->dual_dirs ? p->dual_dirs + (size_t)m * h * d : zero_dirs.data();
        int gs = trajectory_gradient(cx, fs, tj, p->theta, dd, gx.data(), gth.data(), &tcase, &tbest);
        if (gs != ORC_OK && status == ORC_OK) status = gs;
        for (int a = 0; a < d; ++a) out->grad_x[(size_t)m * d + a] = gx[a];
        for (int a = 0; a < nth; ++a) out->grad_theta[(size_t)m * nth + a] = gth[a];
      }
      if (out->best_index) out->best_index[m] = tbest;
      if (out->grad_case) out->grad_case[m] = tcase;
      if (out->status) out->status[m] = status;
      if (out->xs)
        for (int k = 0; k <= h; ++k)
          for (int a = 0; a < d; ++a) out->xs[((size_t)m * (h + 1) + k) * d + a] = fs.X[(size_t)(fs.N + k) * d + a];
      if (out->ys) for (int k = 0; k <= h; ++k) out->ys[(size_t)m * (h + 1) + k] = tj.obs[k];
      if (out->gys)
        for (int k = 0; k <= h; ++k)
          for (int a = 0; a < d; ++a) out->gys[((size_t)m * (h + 1) + k) * d + a] = tj.grads[(size_t)k * d + a];
    }
  }
  return 0;
}

void orc_mean_std(const double* v, int M, int stride, double* mean, double* std_) {
  // Statistics.mean uses pairwise summation; a plain sum in long double is closer to the exact value than either.
  long double s = 0;
  for (int m = 0; m < M; ++m) s += v[(size_t)m * stride];
  double mu = (double)(s / M);
  long double ss = 0;
  for (int m = 0; m < M; ++m) { long double t = v[(size_t)m * stride] - mu; ss += t * t; }
  *mean = mu;
  *std_ = M > 1 ? std::sqrt((double)(ss / (M - 1))) : std::numeric_limits<double>::quiet_NaN();
}

int orc_fit_surrogate(int d, int N, const double* X, int ldX, const double* y, int kernel_id, const double* ktheta,
                      double sigma_n2, double* K, double* L, double* c) {
  Kern kern;
  kern.id = kernel_id;
  for (int i = 0; i < 4; ++i) kern.th[i] = ktheta[i];
  double k0, a, b;
  kern_eval(kern, 0.0, k0, a, b);
  // rbf.jl:161-178
  for (int j = 0; j < N; ++j) {
    K[(size_t)j * N + j] = k0 + sigma_n2;
    for (int i = j + 1; i < N; ++i) {
      double rho2 = 0;
      for (int t = 0; t < d; ++t) { double r = X[(size_t)i * ldX + t] - X[(size_t)j * ldX + t]; rho2 += r * r; }
      double ps;
      kern_eval(kern, std::sqrt(rho2), ps, a, b);
      K[(size_t)j * N + i] = ps;
      K[(size_t)i * N + j] = ps;
    }
  }
  // rbs.jl:93-98 cholesky(Hermitian(K)).L  (column-major output)
  for (size_t i = 0; i < (size_t)N * N; ++i) L[i] = 0;
  for (int j = 0; j < N; ++j) {
    double s = K[(size_t)j * N + j];
    for (int k = 0; k < j; ++k) s -= L[(size_t)k * N + j] * L[(size_t)k * N + j];
    if (!(s > 0)) return 1;
    double ljj = std::sqrt(s);
    L[(size_t)j * N + j] = ljj;
    for (int i = j + 1; i < N; ++i) {
      double t = K[(size_t)j * N + i];
      for (int k = 0; k < j; ++k) t -= L[(size_t)k * N + i] * L[(size_t)k * N + j];
      L[(size_t)j * N + i] = t / ljj;
    }
  }
  // rbs.jl:100-101 c = L' \ (L \ y)
  std::vector<double> Lr((size_t)N * N, 0.0), v(y, y + N);
  for (int i = 0; i < N; ++i)
    for (int k = 0; k <= i; ++k) Lr[(size_t)i * N + k] = L[(size_t)k * N + i];
  fwd_solve(Lr.data(), N, N, v.data(), N, 1);
  bwd_solve(Lr.data(), N, N, v.data(), N, 1);
  for (int i = 0; i < N; ++i) c[i] = v[i];
  return 0;
}

static int build_fs_with_fantasies(const Ctx& cx, FS& fs, int nf, const double* Xf, const double* yf) {
  init_fs(cx.p, fs);
  int st = 0;
  for (int k = 0; k < nf; ++k) {
    int s = condition_fs(cx, fs, Xf + (size_t)k * fs.d, yf[k]);
    if (s && !st) st = s;
  }
  return st;
}

int orc_eval_point(const orc_problem* p, int nf, const double* Xf, const double* yf, const double* x, double* out) {
  Ctx cx;
  setup_ctx(p, cx);
  FS fs;
  int st = build_fs_with_fantasies(cx, fs, nf, Xf, yf);
  SX s;
  eval_fs(cx, fs, x, p->theta, nf - 1, 2, s);
  const int d = p->d;
  double* o = out;
  *o++ = s.mu; *o++ = s.sigma; *o++ = s.g.g; *o++ = s.fstar;
  *o++ = s.g.g_mu; *o++ = s.g.g_sig; *o++ = s.g.g_mumu; *o++ = s.g.g_sigsig;
  *o++ = s.g.g_th; *o++ = s.g.g_thth; *o++ = s.g.g_muth; *o++ = s.g.g_sigth;
  for (int a = 0; a < d; ++a) *o++ = s.dmu[a];
  for (int a = 0; a < d; ++a) *o++ = s.dsig[a];
  for (int a = 0; a < d; ++a) *o++ = s.dal[a];
  for (int i = 0; i < d * d; ++i) *o++ = s.Hmu[i];
  for (int i = 0; i < d * d; ++i) *o++ = s.Hsig[i];
  for (int i = 0; i < d * d; ++i) *o++ = s.Hal_ref[i];
  for (int i = 0; i < d * d; ++i) *o++ = s.Hal_true[i];
  for (int a = 0; a < d; ++a) *o++ = s.d2a_dxdth[a];
  return st;
}

int orc_multistart_solve(const orc_problem* p, int nf, const double* Xf, const double* yf, double* xbest, double* fbest,
                         int* start_status, int* start_iters, double* start_x, double* start_f) {
  Ctx cx;
  setup_ctx(p, cx);
  FS fs;
  build_fs_with_fantasies(cx, fs, nf, Xf, yf);
  int ne;
  return multistart(cx, fs, p->theta, nf - 1, xbest, fbest, &ne, start_status, start_iters, start_x, start_f);
}

// ------------------------------------------------------------------------------------------
// Sobol (utils.jl:4-13 via Sobol.jl, un-pinned): Joe-Kuo direction numbers, Gray-code order,
// origin skipped (Sobol.jl's first point is 0.5,...). 32-bit integers, value = x / 2^32.
// ------------------------------------------------------------------------------------------
static void sobol_dirs(int dim, unsigned V[][32]) {
  for (int j = 0; j < 32; ++j) V[0][j] = 1u << (31 - j);
  for (int dd = 1; dd < dim; ++dd) {
    unsigned pp = orc_sobol_poly[dd];
    int m = 0;
    while ((pp >> (m + 1)) != 0) ++m;
    unsigned long long v[32];
    for (int j = 0; j < m; ++j) v[j] = orc_sobol_minit[dd][j];
    for (int j = m; j < 32; ++j) {
      unsigned long long nv = v[j - m], pow2 = 1;
      for (int k = 0; k < m; ++k) {
        pow2 <<= 1;
        if ((pp >> (m - 1 - k)) & 1u) nv ^= pow2 * v[j - k - 1];
      }
      v[j] = nv;
    }
    for (int j = 0; j < 32; ++j) V[dd][j] = (unsigned)(v[j] << (31 - j));
  }
}

void orc_sobol_uint32(int dim, int npoints, unsigned* out) {
  static unsigned V[ORC_SOBOL_MAXDIM][32];
  sobol_dirs(dim, V);
  std::vector<unsigned> x(dim, 0u);
  for (int k = 1; k <= npoints; ++k) {
    unsigned kk = (unsigned)(k - 1);
    int c = 0;
    while (kk & 1u) { kk >>= 1; ++c; }
    for (int dd = 0; dd < dim; ++dd) {
      x[dd] ^= V[dd][c];
      out[(size_t)(k - 1) * dim + dd] = x[dd];
    }
  }
}

void orc_sobol_uniform(int dim, int npoints, double* out) {
  std::vector<unsigned> u((size_t)dim * npoints);
  orc_sobol_uint32(dim, npoints, u.data());
  for (size_t i = 0; i < u.size(); ++i) out[i] = (double)u[i] / 4294967296.0;
}

void orc_gen_low_discrepancy_sequence(int M, int d, int H, double* out) {
  // utils.jl:65-74
  int offset = ((d + 1) % 2 == 1) ? 1 : 0;
  int D = d + 1 + offset;
  size_t np = (size_t)M * H;
  std::vector<double> S((size_t)D * np), Nn((size_t)D * np);
  orc_sobol_uniform(D, (int)np, S.data());
  // box_muller_transform utils.jl:23-43 (Q8: log10, pair indexing)
  for (size_t j = 0; j < np; ++j) {
    const double* x = S.data() + j * D;
    double* y = Nn.data() + j * D;
    for (int i = 1; i <= D; ++i) {
      if (i % 2 == 1) y[i - 1] = std::sqrt(-2 * std::log10(x[i - 1])) * std::cos(kTwoPi * x[i]);
      else y[i - 1] = std::sqrt(-2 * std::log10(x[i - 2])) * std::sin(kTwoPi * x[i - 1]);
    }
  }
  // reshape(N, M, D, H) column-major, then N[:, 1:end-offset, :]
  const int q = d + 1;
  for (int t = 0; t < H; ++t)
    for (int k = 0; k < q; ++k)
      for (int m = 0; m < M; ++m)
        out[(size_t)m + (size_t)M * k + (size_t)M * q * t] = Nn[(size_t)m + (size_t)M * k + (size_t)M * D * t];
}

void orc_generate_initial_guesses(int S, int d, const double* lbs, const double* ubs, double* out) {
  // utils.jl:145-153
  std::vector<double> u((size_t)d * std::max(S, 1));
  if (S > 0) orc_sobol_uniform(d, S, u.data());
  for (int s = 0; s < S; ++s)
    for (int a = 0; a < d; ++a) out[(size_t)s * d + a] = lbs[a] + (ubs[a] - lbs[a]) * u[(size_t)s * d + a];
  const double eps = 1e-6;
  for (int a = 0; a < d; ++a) out[(size_t)S * d + a] = lbs[a] + eps;
  for (int a = 0; a < d; ++a) out[(size_t)(S + 1) * d + a] = ubs[a] - eps;
}

// decision-rule value and partials (decision_rules.jl:84-127), for the finite-difference ladder:
// out = [g, g_mu, g_sig, g_mumu, g_sigsig, g_muth, g_sigth, g_musig]
void orc_rule_partials(int rule_id, double sigma_tol, double mu, double sigma, double theta1, double fstar, double* out) {
  const GPart g = rule_eval(rule_id, sigma_tol, mu, sigma, theta1, fstar);
  out[0] = g.g; out[1] = g.g_mu; out[2] = g.g_sig; out[3] = g.g_mumu; out[4] = g.g_sigsig; out[5] = g.g_muth; out[6] = g.g_sigth; out[7] = g.g_musig;
}
// the exact trust-region step of the inner solve (solve_tr, optim.jl:9-51): H n x n row-major symmetric; returns hit_constraint
int orc_tr_step(int n, const double* H, const double* g, double Delta, double* p) {
  std::vector<double> A(H, H + (size_t)n * n), work(8 * (size_t)n + 8);
  return tr_step(A.data(), g, n, Delta, p, work.data()) ? 1 : 0;
}
void orc_kernel_scalars(int kernel_id, const double* ktheta, double rho, double* out) {
  Kern k;
  k.id = kernel_id;
  for (int i = 0; i < 4; ++i) k.th[i] = ktheta[i];
  kern_eval(k, rho, out[0], out[1], out[2]);
}

}  // extern "C"
