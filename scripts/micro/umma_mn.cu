// Micro-test: tcgen05.mma kind::tf32, A in tensor memory (TS), B in shared memory in the MN-major no-swizzle layout
// [N / 4 column pieces][32 k rows][4 floats] (the tile-major layout of the saved A / W matrices) - which
// (LBO, SBO, k-step) does the hardware expect?   nvcc -arch=sm_100a -o umma_mn umma_mn.cu && ./umma_mn
#include <cstdio>
#include <vector>
#include <cmath>
#include <cuda_runtime.h>
#include "../../fine_grained_gaussian_process_forcasting_b200/csrc/gpblur_tc.cuh"
using namespace gpblur;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  for (int i = 0; i < 32; i += 4) tc::tmem_st4(taddr + i, make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]));
  tc::tmem_st_wait();
}

template <int N>
__global__ void __launch_bounds__(128, 1) mn_kernel(const float* A, const float* B, float* D, int lbo, int sbo, int kstep,
                                                    int bmajor) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  float* bs = reinterpret_cast<float*>(smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  // B[c][k] (c < N columns, k < 32) -> bs[(c / 4) * 128 + k * 4 + c % 4]
  for (int i = tid; i < N * 32; i += 128) {
    const int c = i / 32, k = i % 32;
    if (bmajor) bs[(c / 4) * 128 + k * 4 + (c % 4)] = B[c * 32 + k];
    else bs[((k / 4) * N + c) * 4 + (k % 4)] = B[c * 32 + k];       // K-major control: [k-chunk][row][4 k]
  }
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_d = slot, tmem_a = slot + 256;
  float v[32];
  for (int k = 0; k < 32; ++k) v[k] = A[tid * 32 + k];
  tmem_st32(tmem_a + ((uint32_t)(warp * 32) << 16), v);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t idesc = tc::make_idesc_tf32(128, N) | ((uint32_t)bmajor << 16);
  if (tid == 0) {
    for (int j = 0; j < 4; ++j) {
      const uint64_t db = tc::make_smem_desc(tc::smem_u32(bs) + j * kstep, lbo, sbo);
      tc::umma_tf32_ts(tmem_d, tmem_a + 8 * j, db, idesc, j ? 1u : 0u);
    }
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
  }
  __syncthreads();
  tc::tc_fence_after();
  for (int c0 = 0; c0 < N; c0 += 32) {
    float o[32];
    tc::tmem_ld32(tmem_d + ((uint32_t)(warp * 32) << 16) + c0, o);
    for (int i = 0; i < 32; ++i) D[tid * N + c0 + i] = o[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(slot, 512);
}

template <int N>
void run(int lbo, int sbo, int kstep, int bmajor) {
  std::vector<float> A(128 * 32), B(N * 32), D(128 * N);
  for (int i = 0; i < 128 * 32; ++i) A[i] = (float)((i * 7 + i / 32) % 13 - 6) * 0.25f;
  for (int i = 0; i < N * 32; ++i) B[i] = (float)((i * 5 + i / 32 * 3) % 11 - 5) * 0.5f;
  float *dA, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = N * 32 * 4 + 1024;
  cudaFuncSetAttribute(mn_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  mn_kernel<N><<<1, 128, smem>>>(dA, dB, dD, lbo, sbo, kstep, bmajor);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < N; ++c) {
      double ref = 0;
      for (int k = 0; k < 32; ++k) ref += (double)A[r * 32 + k] * B[c * 32 + k];
      maxerr = fmax(maxerr, fabs(ref - D[r * N + c]));
    }
  printf("N=%3d bmajor=%d lbo=%4d sbo=%4d kstep=%4d: max |D - ref| = %.3g (%s)\n", N, bmajor, lbo, sbo, kstep, maxerr,
         cudaGetErrorString(e));
  if (bmajor) {
    printf("   D[1][0..7]  :");
    for (int c = 0; c < 8; ++c) printf(" %7.2f", D[1 * N + c]);
    printf("\n   ref[1][0..7]:");
    for (int c = 0; c < 8; ++c) { double ref = 0; for (int k = 0; k < 32; ++k) ref += (double)A[1 * 32 + k] * B[c * 32 + k]; printf(" %7.2f", ref); }
    // which single B column / k-range reproduces D?  brute force: D[1][c] == sum_k A[1][k] * X[k] for X = some k-slice of some column
    printf("\n");
  }
  cudaFree(dA); cudaFree(dB); cudaFree(dD);
}

int main() {
  const int combos[][3] = {{128, 512, 128}, {512, 128, 128}, {128, 512, 256}, {512, 128, 256}, {16, 512, 128}, {512, 16, 128},
                           {128, 64, 128}, {64, 128, 128}, {1024, 512, 128}, {512, 1024, 128}};
  run<64>(64 * 16, 128, 2 * 64 * 16, 0);
  run<256>(256 * 16, 128, 2 * 256 * 16, 0);
  for (int i = 0; i < 4; ++i) run<64>(combos[i][0], combos[i][1], combos[i][2], 1);
  return 0;
}
