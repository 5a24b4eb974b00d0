// Dependent-chain latencies of the fp64 building blocks of the 32 x 32 warp Cholesky (cycles per operation).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f64_lat f64_lat.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(double* out, long long* cyc, double seed) {
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
  const int N = 4096;
  long long t0, t1;
  // DFMA chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = fma(x, y, 1e-12);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
  // DMUL chain
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = x * y;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
  // shuffle of a double
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (i + 1) & 31);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
  // f64 -> f32 -> rsqrtf -> f64
  x = fabs(x) + 1.0;
  t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) x = (double)rsqrtf((float)x) + 1.0;
  t1 = clock64();
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
  // library sqrt + division
  t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < N; ++i) x = 1.0 / sqrt(x + 2.0);
  t1 = clock64();
  if (threadIdx.x == 0) cyc[4] = t1 - t0;
  // independent DFMA throughput: 8 chains per lane, one warp
  double a0 = x, a1 = x + 1, a2 = x + 2, a3 = x + 3, a4 = x + 4, a5 = x + 5, a6 = x + 6, a7 = x + 7;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) {
    a0 = fma(a0, y, 1e-12); a1 = fma(a1, y, 1e-12); a2 = fma(a2, y, 1e-12); a3 = fma(a3, y, 1e-12);
    a4 = fma(a4, y, 1e-12); a5 = fma(a5, y, 1e-12); a6 = fma(a6, y, 1e-12); a7 = fma(a7, y, 1e-12);
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[5] = t1 - t0;
  // shared-memory broadcast round trip: st.shared -> syncwarp -> ld.shared
  __shared__ double buf[64];
  x = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < N; ++i) {
    buf[threadIdx.x & 31] = x;
    __syncwarp();
    x = buf[(i + 1) & 31] + 1.0;
    __syncwarp();
  }
  t1 = clock64();
  if (threadIdx.x == 0) cyc[6] = t1 - t0;
  out[threadIdx.x] = x;
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 32 * 8); cudaMalloc(&cyc, 64);
  for (int rep = 0; rep < 2; ++rep) k<<<1, 32>>>(out, cyc, 1.0);
  long long h[8];
  cudaMemcpy(h, cyc, 56, cudaMemcpyDeviceToHost);
  const char* names[] = {"DFMA dep", "DMUL dep", "shfl f64 dep", "cvt+rsqrtf+cvt+dadd dep", "1/sqrt lib dep", "DFMA x8 indep (per 8)", "st.shared/ld.shared bcast + dadd"};
  for (int i = 0; i < 7; ++i) printf("%-36s %.1f cycles\n", names[i], h[i] / 4096.0);
  return 0;
}
