// latency of dependent DMMA / DFMA chains (single warp) on B200
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(long long* out, double* sink) {
  double c0 = 0, c1 = 0, a = 1.0 + threadIdx.x * 1e-9, b = 1.0;
  long long t0 = clock64();
#pragma unroll
  for (int i = 0; i < 64; ++i) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
  long long t1 = clock64();
  double d0[4] = {0, 0, 0, 0}, d1[4] = {0, 0, 0, 0};
#pragma unroll
  for (int i = 0; i < 64; ++i) asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0[i & 3]), "+d"(d1[i & 3]) : "d"(a), "d"(b));
  long long t2 = clock64();
  double f = a;
#pragma unroll
  for (int i = 0; i < 64; ++i) f = fma(f, b, a);
  long long t3 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; }
  sink[threadIdx.x] = c0 + c1 + d0[0] + d0[1] + d0[2] + d0[3] + d1[0] + d1[1] + d1[2] + d1[3] + f;
}
int main() {
  long long* o; double* s; cudaMalloc(&o, 64); cudaMalloc(&s, 32 * 8);
  for (int r = 0; r < 2; ++r) k<<<1, 32>>>(o, s);
  long long h[3]; cudaMemcpy(h, o, 24, cudaMemcpyDeviceToHost);
  printf("dependent DMMA chain: %.1f cycles each; 4 interleaved chains: %.1f cycles per DMMA; dependent DFMA: %.1f cycles\n", h[0] / 64.0, h[1] / 64.0, h[2] / 64.0);
  return 0;
}
