// Dependent-chain latencies of the operations the per-start logic is made of (one warp, one CTA): DFMA, DADD, double shuffle,
// FP64 division / reciprocal / sqrt / rsqrt, broadcast LDS, __syncwarp. nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o lat_bench lat_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#define N 512
template <class F> __device__ void run(const char* name, double* out, double x, F f, long long* cyc, int id) {
  double v = x;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < N / 8; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v = f(v);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[id] = t1 - t0;
  out[threadIdx.x] += v;
}
__global__ void k(double* out, long long* cyc, double seed, double c1, double c2) {
  __shared__ double sh[64];
  sh[threadIdx.x] = seed + threadIdx.x; sh[threadIdx.x + 32] = c1;
  __syncwarp();
  run("dfma", out, seed, [=](double v) { return fma(v, c1, c2); }, cyc, 0);
  run("dadd", out, seed, [=](double v) { return v + c2; }, cyc, 1);
  run("dmul", out, seed, [=](double v) { return v * c1; }, cyc, 2);
  run("shfl", out, seed, [=](double v) { return __shfl_xor_sync(0xffffffffu, v, 1); }, cyc, 3);
  run("shfl+add", out, seed, [=](double v) { return v + __shfl_xor_sync(0xffffffffu, v, 2); }, cyc, 4);
  run("div", out, seed, [=](double v) { return c1 / v + c2; }, cyc, 5);
  run("rcp", out, seed, [=](double v) { return __drcp_rn(v) + c2; }, cyc, 6);
  run("sqrt", out, seed, [=](double v) { return sqrt(v) + c2; }, cyc, 7);
  run("rsqrt", out, seed, [=](double v) { return rsqrt(v) + c2; }, cyc, 8);
  run("lds", out, seed, [&](double v) { return sh[((int)v) & 31] ; }, cyc, 9);
  run("lds+fma", out, seed, [&](double v) { return fma(sh[32 + (((int)__double2hiint(v)) & 1)], v, c2); }, cyc, 10);
  run("syncwarp", out, seed, [&](double v) { __syncwarp(); return v; }, cyc, 11);
  run("ffma32", out, seed, [=](double v) { return (double)fmaf((float)v, 1.0001f, 0.5f); }, cyc, 12);
  run("fast rcp (mufu + 2 newton)", out, seed, [=](double v) { double r; asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(v)); double e = fma(-v, r, 1.0); r = fma(r, e, r); e = fma(-v, r, 1.0); r = fma(r, e, r); return r + c2; }, cyc, 13);
  run("log", out, seed, [=](double v) { return log(v) + c2; }, cyc, 14);
  run("exp", out, seed, [=](double v) { return exp(-v) + c2; }, cyc, 15);
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMemset(out, 0, 32 * 8); cudaMalloc(&cyc, 16 * 8);
  const char* names[16] = {"DFMA", "DADD", "DMUL", "shfl (double)", "shfl + DADD", "c / v + c (IEEE division)", "__drcp_rn + DADD", "sqrt + DADD", "rsqrt + DADD", "LDS (dependent address)", "LDS + DFMA", "__syncwarp", "cvt + FFMA + cvt", "rcp.approx + 2 Newton + DADD", "log + DADD", "exp + DADD"};
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(out, cyc, 1.25, 1.0000001, 0.75);
  long long h[16]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaDeviceSynchronize();
  printf("dependent-chain latency, one warp (B200), cycles per operation (%s)\n", cudaGetErrorString(e));
  for (int i = 0; i < 16; ++i) printf("  %-34s %7.1f\n", names[i], (double)h[i] / N);
  return 0;
}
