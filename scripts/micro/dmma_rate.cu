// FP64 mma.sync m8n8k4 issue rate / latency on one SM (8 warps).  GPU box only.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int CH>
__global__ void k(double* out, long long* cyc) {
  double c[CH][2];
  for (int i = 0; i < CH; ++i) { c[i][0] = threadIdx.x; c[i][1] = 1.0; }
  const double a = 1.0000001, b = 0.9999999;
  __syncthreads();
  const long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < 1024; ++it)
#pragma unroll
    for (int i = 0; i < CH; ++i) dmma884(c[i][0], c[i][1], a, b);
  const long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < CH; ++i) s += c[i][0] + c[i][1];
  out[threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
  double* out; long long* cyc; long long h;
  cudaMalloc(&out, 8 * 1024); cudaMalloc(&cyc, 8);
  for (int warps = 1; warps <= 8; warps *= 2) {
    k<1><<<1, 32 * warps>>>(out, cyc); k<1><<<1, 32 * warps>>>(out, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps %d chains 1: %.1f cycles per dependent dmma\n", warps, h / 1024.0);
    k<4><<<1, 32 * warps>>>(out, cyc); k<4><<<1, 32 * warps>>>(out, cyc);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("warps %d chains 4: %.1f cycles per dmma per warp, %.2f cycles per dmma per SM\n", warps, h / 4096.0, h / 4096.0 / warps);
  }
  return 0;
}
