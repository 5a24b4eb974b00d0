// Micro-test for the next step named in DESIGN.md: tcgen05.mma.kind::tf32 with the A operand in TENSOR MEMORY
// ("TS" form), written by the threads that own the rows (tcgen05.st.32x32b: lane = row, registers = K columns).
// Checks D = A B^T (M = 128, N in {64, 256}, K = 32) against the host and times the MMA rate next to the SS form.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_ts umma_ts.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../fine_grained_gaussian_process_forcasting_b200/csrc/gpblur_tc.cuh"
using namespace gpblur;

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16,"
      " %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};\n"
      :
      : "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
        "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
        "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
        "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
        "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])),
        "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])),
        "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])),
        "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])),
        "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// A [128][32] row-major, B [N][32] row-major (both exactly representable in TF32) -> D [128][N] = A B^T
template <int N>
__global__ void __launch_bounds__(128, 1) ts_kernel(const float* A, const float* B, float* D, long long* cyc, int reps) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t slot;
  float* bs = reinterpret_cast<float*>(smem);               // B plane: [8 chunks][N rows][4]
  float* as = bs + 8 * N * 4;                               // A plane for the SS timing run
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < N * 8; i += 128) {
    const int r = i % N, c = i / N;
    *reinterpret_cast<float4*>(bs + (c * N + r) * 4) = *reinterpret_cast<const float4*>(B + r * 32 + c * 4);
  }
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<float4*>(as + (c * 128 + tid) * 4) = *reinterpret_cast<const float4*>(A + tid * 32 + c * 4);
  if (warp == 0) tc::tmem_alloc(&slot, 512);
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_d = slot, tmem_a = slot + 256;
  // every thread owns row tid: its 32 K-values go to TMEM lane tid, columns [256, 288)
  float v[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = A[tid * 32 + k];
  tmem_st32(tmem_a + ((uint32_t)(warp * 32) << 16), v);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t idesc = tc::make_idesc_tf32(128, N);
  if (tid == 0) {
    for (int j = 0; j < 4; ++j) {
      const uint64_t db = tc::make_smem_desc(tc::smem_u32(bs) + 2 * j * N * 16, N * 16, 128);
      umma_tf32_ts(tmem_d, tmem_a + 8 * j, db, idesc, j ? 1u : 0u);
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
  tc::tc_fence_after();
  // ---- rate: TS vs SS, `reps` slabs of 4 MMAs ----
  if (tid == 0) {
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int j = 0; j < 4; ++j) {
        const uint64_t db = tc::make_smem_desc(tc::smem_u32(bs) + 2 * j * N * 16, N * 16, 128);
        umma_tf32_ts(tmem_d, tmem_a + 8 * j, db, idesc, 1u);
      }
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 1);
    long long t1 = clock64();
    for (int r = 0; r < reps; ++r)
      for (int j = 0; j < 4; ++j) {
        const uint64_t db = tc::make_smem_desc(tc::smem_u32(bs) + 2 * j * N * 16, N * 16, 128);
        const uint64_t da = tc::make_smem_desc(tc::smem_u32(as) + 2 * j * 128 * 16, 128 * 16, 128);
        tc::umma_tf32(tmem_d, da, db, idesc, 1u);
      }
    tc::umma_commit(&bar);
    tc::mbar_wait(&bar, 0);
    long long t2 = clock64();
    cyc[0] = t1 - t0;
    cyc[1] = t2 - t1;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(slot, 512);
}

template <int N>
void run() {
  std::vector<float> A(128 * 32), B(N * 32), D(128 * N);
  for (int i = 0; i < 128 * 32; ++i) A[i] = (float)((i * 7 + i / 32) % 13 - 6) * 0.25f;
  for (int i = 0; i < N * 32; ++i) B[i] = (float)((i * 5 + i / 32 * 3) % 11 - 5) * 0.5f;
  float *dA, *dB, *dD;
  long long* dc;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dc, 16);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = (8 * N * 4 + 8 * 128 * 4) * 4 + 1024;
  cudaFuncSetAttribute(ts_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int reps = 500;
  ts_kernel<N><<<148, 128, smem>>>(dA, dB, dD, dc, reps);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  long long h[2];
  cudaMemcpy(h, dc, 16, cudaMemcpyDeviceToHost);
  double maxerr = 0;
  for (int r = 0; r < 128; ++r)
    for (int c = 0; c < N; ++c) {
      double ref = 0;
      for (int k = 0; k < 32; ++k) ref += (double)A[r * 32 + k] * B[c * 32 + k];
      maxerr = fmax(maxerr, fabs(ref - D[r * N + c]));
    }
  printf("TS-mode tf32 M=128 N=%3d K=32: max |D - ref| = %.3g (%s)   cycles per MMA: TS %.1f, SS %.1f\n", N, maxerr,
         cudaGetErrorString(e), (double)h[0] / (4.0 * reps), (double)h[1] / (4.0 * reps));
}

int main() {
  run<64>();
  run<128>();
  run<256>();
  return 0;
}
