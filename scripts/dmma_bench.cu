// Micro-benchmark: FP64 vector FMA vs DMMA (mma.sync.m8n8k4.f64) throughput on one GPU.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_dfma(double* out, int iters) {
  double a[16];
  for (int i = 0; i < 16; ++i) a[i] = 1.0 + 1e-9 * (threadIdx.x + i);
  const double b = 1.0000001, c = 1e-9;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fma(a[i], b, c);
  double s = 0; for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_dmma(double* out, int iters) {
  double c[8][2];
  for (int i = 0; i < 8; ++i) { c[i][0] = 0.0; c[i][1] = 0.0; }
  double a = 1.0 + 1e-9 * threadIdx.x, b = 1.0 - 1e-9 * threadIdx.x;
  for (int it = 0; it < iters; ++it)
#pragma unroll
    for (int i = 0; i < 8; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
  double s = 0; for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 4 * 512);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {128, 256, 512}) {
    for (int which = 0; which < 2; ++which) {
      float best = 1e30f; int iters = 4096;
      for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) k_dfma<<<sms * 4, threads>>>(out, iters); else k_dmma<<<sms * 4, threads>>>(out, iters);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = ms < best ? ms : best;
      }
      double fl = which == 0 ? (double)sms * 4 * threads * iters * 16 * 2 : (double)sms * 4 * (threads / 32) * iters * 8 * 256 * 2;
      printf("%s threads/block=%d blocks=%d: %.2f TFLOP/s (%.3f ms)\n", which == 0 ? "DFMA" : "DMMA m8n8k4", threads, sms * 4, fl / (best * 1e-3) / 1e12, best);
    }
  }
  return 0;
}
