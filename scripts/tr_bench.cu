// Stand-alone timing of the per-start trust-region step (one warp): first call (cold instruction cache) against repeated calls
// (warm), for the register-resident path and the general shared-memory path. Includes the kernel translation unit to reach the
// functions in its anonymous namespace.  nvcc -O3 -std=c++17 --expt-relaxed-constexpr --extended-lambda -gencode arch=compute_100a,code=sm_100a -I../rollout-bayesian-optimization_b200/csrc -I../include -o tr_bench tr_bench.cu
#include "../rollout-bayesian-optimization_b200/csrc/rollout_kernel.cu"
#include <cstdio>
#include <cstdlib>
namespace rbo {
__global__ void __launch_bounds__(512, 1) tr_bench_kernel(const double* Hin, const double* gin, int n, double Delta, long long* cyc, double* pout, int which, int reps) {
  __shared__ double H[256], g[16], A[256], ta[16], te[16], yv[16];
  __shared__ int fr[32];
  const int lane = threadIdx.x;
  for (int i = lane; i < n * n; i += 32) H[i] = Hin[i];
  if (lane < n) { g[lane] = gin[lane]; fr[lane] = lane; }
  __syncwarp();
  for (int r = 0; r < reps; ++r) {
    long long t0 = clock64();
    bool hit;
    if (which == 0) hit = tr_step16(H, g, fr, n, n, Delta, ta, te, yv);
    else hit = tr_step_warp(H, g, fr, n, n, Delta, A, yv);
    long long t1 = clock64();
    if (lane == 0) cyc[r] = t1 - t0;
    if (lane < n) pout[which * 16 + lane] = yv[lane] + (hit ? 0.0 : 1e-300);
    __syncwarp();
  }
}
}  // namespace rbo
int main() {
  const int n = 10, reps = 16;
  double H[n * n], g[n];
  srand(3);
  for (int i = 0; i < n; ++i) { g[i] = rand() / (double)RAND_MAX - 0.5; for (int j = 0; j <= i; ++j) { double v = rand() / (double)RAND_MAX - 0.5; H[i * n + j] = v; H[j * n + i] = v; } }
  double *dH, *dg, *dp; long long* dc;
  cudaMalloc(&dH, sizeof(H)); cudaMalloc(&dg, sizeof(g)); cudaMalloc(&dp, 32 * 8); cudaMalloc(&dc, reps * 8);
  cudaMemcpy(dH, H, sizeof(H), cudaMemcpyHostToDevice); cudaMemcpy(dg, g, sizeof(g), cudaMemcpyHostToDevice);
  for (int which = 0; which < 2; ++which) {
    rbo::tr_bench_kernel<<<1, 32>>>(dH, dg, n, 0.3, dc, dp, which, reps);
    long long c[reps]; double p[32];
    cudaMemcpy(c, dc, sizeof(c), cudaMemcpyDeviceToHost); cudaMemcpy(p, dp, sizeof(p), cudaMemcpyDeviceToHost);
    printf("%s n=%d (indefinite H, boundary step): cycles per call:", which == 0 ? "tr_step16 (registers)" : "tr_step_warp (shared memory)", n);
    for (int r = 0; r < reps; ++r) printf(" %lld", c[r]);
    double nn = 0; for (int i = 0; i < n; ++i) nn += p[which * 16 + i] * p[which * 16 + i];
    printf("\n   |p| = %.12f  p[0..2] = %.12f %.12f %.12f  (%s)\n", sqrt(nn), p[which * 16], p[which * 16 + 1], p[which * 16 + 2], cudaGetErrorString(cudaGetLastError()));
  }
#ifdef RBO_PHASE_TIMERS
  unsigned long long t[16];
  cudaMemcpyFromSymbol(t, rbo::g_tr_cycles, sizeof(t));
  const char* nm[10] = {"", "", "", "", "load", "householder", "write-out + gershgorin", "probes", "tridiagonal solve", "back-transformation"};
  for (int i = 4; i < 10; ++i) printf("  tr_step16 %-24s %8.0f cycles per call (avg over %d calls incl. the cold one)\n", nm[i], (double)t[i] / reps, reps);
#endif
  return 0;
}
