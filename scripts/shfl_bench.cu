// Throughput of independent warp shuffles / shared-memory loads issued by ONE warp (B200): cycles per instruction.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float* out, long long* cyc, int src) {
  __shared__ double sh[512];
  const int lane = threadIdx.x;
  for (int i = lane; i < 512; i += 32) sh[i] = i;
  __syncwarp();
  float x[8];
  for (int j = 0; j < 8; ++j) x[j] = lane * 8 + j;
  long long t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = __shfl_sync(0xffffffffu, x[j], src + j);   // 8 independent 32-bit shuffles (register lane index)
  }
  long long t1 = clock64();
  if (lane == 0) cyc[0] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = __shfl_xor_sync(0xffffffffu, x[j], 1 + (j & 3));   // butterfly, immediate
  }
  t1 = clock64();
  if (lane == 0) cyc[1] = t1 - t0;
  double y[8];
  for (int j = 0; j < 8; ++j) y[j] = x[j];
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] += sh[(src + 8 * i + j) & 511];     // 8 independent broadcast LDS.64 + DADD
  }
  t1 = clock64();
  if (lane == 0) cyc[2] = t1 - t0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) y[j] += sh[(lane + src + 8 * i + j) & 511];   // 8 independent per-lane LDS.64 + DADD
  }
  t1 = clock64();
  if (lane == 0) cyc[3] = t1 - t0;
  // one dependent shuffle chain
  float z = lane;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) z = __shfl_sync(0xffffffffu, z, (int)z + src + 1) + 1.0f;
  }
  t1 = clock64();
  if (lane == 0) cyc[4] = t1 - t0;
  float s = z; for (int j = 0; j < 8; ++j) s += x[j] + (float)y[j];
  out[lane] = s;
}
int main() {
  float* out; long long* cyc; cudaMalloc(&out, 128); cudaMalloc(&cyc, 64);
  for (int r = 0; r < 2; ++r) k<<<1, 32>>>(out, cyc, 1);
  long long h[5]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  printf("one warp, 512 instructions each (%s)\n", cudaGetErrorString(cudaDeviceSynchronize()));
  const char* nm[5] = {"independent SHFL.IDX (32-bit)", "independent SHFL.BFLY (32-bit)", "independent broadcast LDS.64 + DADD", "independent per-lane LDS.64 + DADD", "dependent SHFL.IDX + FADD chain"};
  for (int i = 0; i < 5; ++i) printf("  %-40s %6.2f cycles per instruction\n", nm[i], (double)h[i] / 512);
  return 0;
}
