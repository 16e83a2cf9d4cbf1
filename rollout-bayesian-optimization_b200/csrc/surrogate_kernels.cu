// surrogate_kernels.cu -- device side of rbo_set_surrogate / rbo_condition: everything that turns the caller's base
// surrogate (X, L, y, c: what FantasySurrogate(s, h) copies, rbs.jl:345-381) into the resident, packed form the rollout
// kernel streams, and the rank-1 extension of that form by one observation (condition!(::Surrogate), rbs.jl:214-222 ->
// update_covariance! :166-183, update_cholesky! :185-203, update_coefficients! :205-212), so that the factor never leaves the
// GPU between Bayesian-optimisation iterations.
//
//   rbo_trinv_kernel        explicit inverse of the lower Cholesky factor, one warp per column, forward substitution in
//                           double-double arithmetic (one rounding per stored entry, like the extended-precision host loop
//                           it replaces: O(n^3) on the host was 0.5 s at n = 1000)
//   rbo_pack_*_kernel       L0^-1 -> forward / backward 32-row panels (k-major, pitch RBO_LP, 32-k chunks) and the backward
//                           panels once more in mma.m8n8k4 A-fragment order
//   rbo_matvec_*            u0 = L0^-1 y, c0 = L0^-T u0, the new factor row l = L0^-1 k and the new row of the inverse
#include "rbo_kernel.cuh"

namespace rbo {

namespace {
__device__ __forceinline__ void two_sum(double a, double b, double& s, double& e) {
  s = a + b;
  const double bb = s - a;
  e = (a - (s - bb)) + (b - bb);
}
// (bh, bl) -= l * (xh, xl)
__device__ __forceinline__ void dd_sub_prod(double& bh, double& bl, double l, double xh, double xl) {
  const double ph = l * xh, pe = fma(l, xh, -ph) + l * xl;
  double s, e;
  two_sum(bh, -ph, s, e);
  e += bl - pe;
  bh = s + e;
  bl = e - (bh - s);
}
// (xh, xl) = (bh, bl) / l
__device__ __forceinline__ void dd_div(double bh, double bl, double l, double& xh, double& xl) {
  const double q1 = bh / l;
  const double r = fma(-q1, l, bh) + bl;
  const double q2 = r / l;
  xh = q1 + q2;
  xl = q2 - (xh - q1);
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}
}  // namespace

// Column j of Linv = L^-1 (row-major, pitch ldi, zero above the diagonal, identity on the padding rows/columns N..N32-1).
// L: column-major N x N (ld = N). One warp per column; its right-hand side lives in shared memory as double-double.
__global__ void rbo_trinv_kernel(const double* __restrict__ L, int N, int N32, double* __restrict__ Linv, int ldi) {
  extern __shared__ double sm_trinv[];
  const int wpc = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j = blockIdx.x * wpc + warp;
  if (j >= N32) return;
  if (j >= N) {
    if (lane == 0) Linv[(size_t)j * ldi + j] = 1.0;
    return;
  }
  double* bh = sm_trinv + (size_t)warp * 2 * N;
  double* bl = bh + N;
  for (int i = j + lane; i < N; i += 32) { bh[i] = (i == j) ? 1.0 : 0.0; bl[i] = 0.0; }
  __syncwarp();
  for (int k = j; k < N; ++k) {
    const double* Lk = L + (size_t)k * N;
    double xh, xl;
    dd_div(bh[k], bl[k], Lk[k], xh, xl);
    if (lane == 0) Linv[(size_t)k * ldi + j] = xh + xl;
    for (int i = k + 1 + lane; i < N; i += 32) dd_sub_prod(bh[i], bl[i], Lk[i], xh, xl);
    __syncwarp();
  }
}

// forward panels: panel ib (rows r0 = 32 ib ..), k = 0 .. r0 + 31: Lf[base(ib) + k * LP + r] = (k <= r0 + r) ? Linv[r0 + r][k] : 0
__global__ void rbo_pack_fwd_kernel(const double* __restrict__ Linv, int ldi, int nb32, double* __restrict__ Lf) {
  const int ib = blockIdx.y, r0 = RBO_BR * ib, nk = r0 + RBO_BR;
  double* pf = Lf + (size_t)RBO_LP * RBO_BR * ((size_t)ib * (ib + 1) / 2);
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nk * RBO_BR; idx += gridDim.x * blockDim.x) {
    const int k = idx / RBO_BR, r = idx - k * RBO_BR;
    pf[(size_t)k * RBO_LP + r] = (k <= r0 + r) ? Linv[(size_t)(r0 + r) * ldi + k] : 0.0;
  }
}
// backward panels: panel ib: kk = 0 .. N32 - r0 - 33: Linv[r0 + 32 + kk][r0 + r]; then the transposed diagonal block kk = 0..31:
// Linv[r0 + kk][r0 + r] (kk >= r). Lbf: the same panel in mma.m8n8k4 A-fragment order (chunk -> 4 row-quarter tiles of 8 rows x
// 32 k -> 4 k-pairs -> lane (g = row, tg) -> {k = 8 p + tg, 8 p + 4 + tg}).
__global__ void rbo_pack_bwd_kernel(const double* __restrict__ Linv, int ldi, int nb32, double* __restrict__ Lb, double* __restrict__ Lbf) {
  const int ib = blockIdx.y, r0 = RBO_BR * ib, N32 = nb32 * RBO_BR, k0 = r0 + RBO_BR, nkb = N32 - k0, nk = nkb + RBO_BR;
  const size_t chunk0 = (size_t)nb32 * ib - (size_t)ib * (ib - 1) / 2;
  double* pb = Lb + (size_t)RBO_LP * RBO_BR * chunk0;
  double* pf = Lbf + (size_t)1024 * chunk0;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < nk * RBO_BR; idx += gridDim.x * blockDim.x) {
    const int kk = idx / RBO_BR, r = idx - kk * RBO_BR;
    double v;
    if (kk < nkb) v = Linv[(size_t)(k0 + kk) * ldi + r0 + r];
    else { const int k2 = kk - nkb; v = (k2 >= r) ? Linv[(size_t)(r0 + k2) * ldi + r0 + r] : 0.0; }
    pb[(size_t)kk * RBO_LP + r] = v;
    // A-fragment position of (k = kk within its chunk, row r)
    const int cc = kk / RBO_BR, kc = kk - cc * RBO_BR, rq = r >> 3, g = r & 7, pp = kc >> 3, e = (kc >> 2) & 1, tg = kc & 3;
    pf[(((size_t)cc * 4 + rq) * 4 + pp) * 64 + 2 * (g * 4 + tg) + e] = v;
  }
}

// out[i] = sum_{j <= i} Linv[i][j] v[j] (i < n): one warp per row
__global__ void rbo_matvec_lower_kernel(const double* __restrict__ Linv, int ldi, int n, const double* __restrict__ v, double* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  double acc = 0.0;
  for (int j = lane; j <= row; j += 32) acc = fma(Linv[(size_t)row * ldi + j], v[j], acc);
  acc = warp_sum_d(acc);
  if (lane == 0) out[row] = acc;
}
// out[j] = scale * sum_{i >= j, i < n} Linv[i][j] v[i] (j < n): thread per column, coalesced across columns
__global__ void rbo_matvec_lower_t_kernel(const double* __restrict__ Linv, int ldi, int n, const double* __restrict__ v, double scale, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  double acc = 0.0;
  for (int i = j; i < n; ++i) acc = fma(Linv[(size_t)i * ldi + j], v[i], acc);
  out[j] = scale * acc;
}

// Xb[p][j] (coordinate-major, pitch N8, pad columns 0) from the point-major resident copy Xpts[j][p]
__global__ void rbo_layout_x_kernel(const double* __restrict__ Xpts, int d, int N, int N8, double* __restrict__ Xb) {
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < d * N8; idx += gridDim.x * blockDim.x) {
    const int p = idx / N8, j = idx - p * N8;
    Xb[idx] = (j < N) ? Xpts[(size_t)j * d + p] : 0.0;
  }
}
// k[j] = psi(|x - X_j|) for the N resident points (update_covariance!, rbs.jl:166-183 / rbf.jl:180-191)
__global__ void rbo_kvec_kernel(const double* __restrict__ Xpts, int d, int N, const double* __restrict__ x, KernelSpec kern, double* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= N) return;
  double rho2 = 0.0;
  for (int p = 0; p < d; ++p) { const double r = x[p] - Xpts[(size_t)j * d + p]; rho2 = fma(r, r, rho2); }
  double psi, a, b;
  kern_eval(kern, sqrt(rho2), psi, a, b);
  out[j] = psi;
}
// scal[0] = l.l, scal[1] = l.u (one warp)
__global__ void rbo_dots_kernel(const double* __restrict__ l, const double* __restrict__ u, int n, double* __restrict__ scal) {
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < n; i += 32) { a = fma(l[i], l[i], a); b = fma(l[i], u[i], b); }
  a = warp_sum_d(a); b = warp_sum_d(b);
  if (threadIdx.x == 0) { scal[0] = a; scal[1] = b; }
}
// The new (n-th) row of the inverse and of u = L^-1 y once l = L^-1 k is known (update_cholesky!, rbs.jl:185-203):
//   l_nn = sqrt(k0 + sigma_n2 - l.l);  Linv[n][j] = -(l' Linv)_j / l_nn (tmp holds l' Linv), Linv[n][n] = 1 / l_nn;
//   u[n] = (y_new - l.u) / l_nn.  status[0] = 1 if the pivot is not positive (PosDefException in the reference).
__global__ void rbo_append_row_kernel(double* __restrict__ Linv, int ldi, int n, const double* __restrict__ tmp, const double* __restrict__ scal,
                                      double kdiag, double ynew, double* __restrict__ u, int* __restrict__ status) {
  const double s = kdiag - scal[0];
  const bool ok = s > 0.0;
  const double lnn = sqrt(s), inv = 1.0 / lnn;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j <= n; j += gridDim.x * blockDim.x) {
    if (j < n) Linv[(size_t)n * ldi + j] = ok ? -tmp[j] * inv : 0.0;
    else {
      Linv[(size_t)n * ldi + n] = ok ? inv : 1.0;
      u[n] = ok ? (ynew - scal[1]) * inv : 0.0;
      status[0] = ok ? 0 : 1;
    }
  }
}

}  // namespace rbo
