// Micro-benchmark: cycles per tcgen05.mma.kind::tf32 (M=128, N in {32..256}, K=8), SS operands in the K-major
// no-swizzle layout used by gpblur, issued back to back by one thread.  nvcc -arch=sm_100a -o umma_rate umma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "../../fine_grained_gaussian_process_forcasting_b200/csrc/gpblur_tc.cuh"
using namespace gpblur;

template <int N, int KIND>   // KIND 0: tf32, 1: f16 (bf16 inputs)
__global__ void __launch_bounds__(128, 1) rate_kernel(long long* out, int reps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  float* a = reinterpret_cast<float*>(smem);
  float* b = a + 8 * 128 * 4;
  for (int i = threadIdx.x; i < 8 * 128 * 4 + 8 * 256 * 4; i += blockDim.x) a[i] = 1.0f;
  if (threadIdx.x < 32) tc::tmem_alloc(&slot, 512);
  if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t d = slot;
    uint32_t idesc = tc::make_idesc_tf32(128, N);
    if (KIND == 1) idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t da = tc::make_smem_desc(tc::smem_u32(a), 128 * 16, 128);
    const uint64_t db = tc::make_smem_desc(tc::smem_u32(b), N * 16, 128);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (KIND == 0) tc::umma_tf32(d, da, db, idesc, r ? 1u : 0u);
      else asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(d),
                        "l"(da), "l"(db), "r"(idesc), "r"(r ? 1u : 0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
    }
    long long t1 = clock64();
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tc::tmem_dealloc(slot, 512);
}

template <int N, int KIND>
void run(const char* name) {
  long long* d;
  cudaMalloc(&d, 16);
  const int smem = (8 * 128 * 4 + 8 * 256 * 4) * 4 + 1024;
  cudaFuncSetAttribute(rate_kernel<N, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 2000;
  for (int it = 0; it < 2; ++it) rate_kernel<N, KIND><<<148, 128, smem>>>(d, reps);
  cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%s N=%3d: issue %.1f cyc/MMA, complete %.1f cyc/MMA  (%s)\n", name, N, (double)h[0] / reps, (double)h[1] / reps,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}

int main() {
  run<256, 0>("tf32"); run<128, 0>("tf32"); run<64, 0>("tf32"); run<32, 0>("tf32");
  run<256, 1>("f16 "); run<128, 1>("f16 ");
  return 0;
}
