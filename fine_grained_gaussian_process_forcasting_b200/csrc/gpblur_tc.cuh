// tcgen05 (5th-gen tensor core) building blocks for the 3xTF32 GEMMs: TMEM allocation, mbarrier, UMMA shared
// memory / instruction descriptors, the error-compensated operand split and TMEM -> register loads.
//
// Operand tiles are written by CUDA cores (they have to be: every fp32 operand is split into a TF32 "hi" and a
// TF32 "lo" plane, D += Ahi*Bhi + Ahi*Blo + Alo*Bhi), in the canonical K-major NO-SWIZZLE UMMA layout
//   byte offset(row r, 16-byte k-chunk c) = c * (ROWS * 16) + r * 16
// i.e. per k-chunk (4 fp32) a dense [ROWS][16 B] plane: core matrices (8 rows x 16 B) are contiguous along the
// rows (SBO = 128 B) and ROWS*16 B apart along K (LBO).  One tcgen05.mma.kind::tf32 consumes K = 8 = two chunks.
#pragma once
#include "gpblur_common.cuh"

namespace gpblur {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::); }
// wait for the phase parity; traps instead of hanging the GPU if the phase never completes.  The suspend-time hint
// lets the hardware park the thread until the phase completes (or the hint expires) instead of re-polling: in the
// point kernels ~25 % of all issued instructions were poll-loop overhead competing with the other warps' work
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
  uint32_t ok = 0;
  for (uint32_t tries = 0; !ok; ++tries) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"(20000u)
        : "memory");
    if (tries > (1u << 22)) asm volatile("trap;");
  }
}

// non-blocking probe of the phase parity
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}

// transaction barrier for cp.async.bulk: one arrival + `bytes` expected
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D TMA bulk copy global -> shared (size multiple of 16 B), completion signalled on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// 1-D TMA bulk copy shared -> global (bulk async-group completion): the epilogues stage a row segment in shared
// memory and let the TMA engine write it out as full lines instead of 32 scattered 16-byte stores per instruction
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, const void* src_smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst_gmem), "r"(smem_u32(src_smem)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// every bulk store committed by this thread has finished READING its shared-memory source
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
// value of lane 0, which also tells the compiler that the result is warp-uniform (uniform registers, no
// per-thread -> uniform conversion loops around the tcgen05 instructions)
__device__ __forceinline__ uint32_t uniform_u32(uint32_t v) { return __shfl_sync(0xffffffffu, v, 0); }
__device__ __forceinline__ uint64_t uniform_u64(uint64_t v) {
  return ((uint64_t)__shfl_sync(0xffffffffu, (uint32_t)(v >> 32), 0) << 32) | __shfl_sync(0xffffffffu, (uint32_t)v, 0);
}

// ---- fences ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM -----------------------------------------------------------------------------------------------
// one full warp allocates `ncols` (power of two >= 32) columns; the base address lands in *slot (shared)
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(ncols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
}
__device__ __forceinline__ void tmem_dealloc(uint32_t base, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(ncols));
}
// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (lane = row of the tile)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15,"
      " %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// split form of tmem_ld16: issue the load now, wait later (the registers carry a dependency through the wait), so
// that the TMEM round trip overlaps the work on the previous chunk
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32"
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16_wait(uint32_t (&r)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// ---- A operand in tensor memory (TS form) ---------------------------------------------------------------
// For kind::tf32 with M = 128 the A tile lives at TMEM lane = row, column = base + k (one fp32 cell per element;
// verified by scripts/micro/umma_ts.cu).  A thread that owns row `lane` of its warp's lane quadrant writes 4
// consecutive k-values with one tcgen05.st; the MMA then reads A from TMEM and only B from shared memory.
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const float4& v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(__float_as_uint(v.x)),
               "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w))
               : "memory");
}
// 16 consecutive k-values of the thread's row (TMEM lane) in one instruction
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// ---- descriptors ----------------------------------------------------------------------------------------
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 | [32,46) stride byte offset >> 4 |
//   [46,48) version = 1 (sm_100) | [61,64) layout type = 0
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor for kind::tf32, fp32 accumulate, both operands K-major (cute::UMMA::InstrDescriptor):
//   [4,6) c_format = 1 (F32) | [7,10) a_format = 2 (TF32) | [10,13) b_format = 2 | [17,23) N >> 3 | [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  // form without the disable-output-lane vector: fewer operands for the single issuing thread to set up
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on `bar` when every tcgen05.mma issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// exp(x) for x <= 0 via ex2.approx with a compensated x * log2(e) (relative error ~2^-22, ~5 instructions)
__device__ __forceinline__ float fast_exp(float x) {
  const float t = x * 1.4426950408889634f;
  const float r = fmaf(x, 1.4426950408889634f, -t);            // rounding error of the product
  const float t2 = fmaf(x, 1.9259629911266175e-08f, r) + t;     // + x * (log2e - fl(log2e))
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t2));
  return y;
}

// 2^x, one MUFU (relative error ~2^-22); the exponent arrives pre-scaled by log2(e)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---- 3xTF32 split ---------------------------------------------------------------------------------------
// hi = round-to-nearest TF32 of a (low 13 mantissa bits zero), lo = a - hi (exact in fp32; the tensor core
// truncates it to TF32 at 2^-22 |a|).  a*b ~= hi_a*hi_b + hi_a*lo_b + lo_a*hi_b, relative error ~2^-21.
__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
  uint32_t h;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(a));
  hi = __uint_as_float(h);
  lo = a - hi;
}
__device__ __forceinline__ void store_split(float* hi_plane, float* lo_plane, int off_floats, const float4& v) {
  float4 h, l;
  split_tf32(v.x, h.x, l.x); split_tf32(v.y, h.y, l.y); split_tf32(v.z, h.z, l.z); split_tf32(v.w, h.w, l.w);
  *reinterpret_cast<float4*>(hi_plane + off_floats) = h;
  *reinterpret_cast<float4*>(lo_plane + off_floats) = l;
}

// byte layout helper: float offset of (row r, k-chunk c) in a [KCH chunks][ROWS][4 floats] operand plane
template <int ROWS>
__device__ __forceinline__ int op_off(int r, int c) { return (c * ROWS + r) * 4; }

// issue the 3 x (KT / 8) MMAs of one K slab: A planes [KT/4 chunks][128 rows], B planes [KT/4][NROWS]
template <int KT, int NROWS>
__device__ __forceinline__ void issue_slab_3xtf32(uint32_t tmem_d, const float* a_hi, const float* a_lo,
                                                  const float* b_hi, const float* b_lo, uint32_t idesc, bool first,
                                                  int b_rows = NROWS) {
  constexpr uint32_t A_LBO = 128 * 16, SBO = 128;
  const uint32_t B_LBO = (uint32_t)b_rows * 16;
  const uint32_t ah = smem_u32(a_hi), al = smem_u32(a_lo), bh = smem_u32(b_hi), bl = smem_u32(b_lo);
#pragma unroll
  for (int j = 0; j < KT / 8; ++j) {
    const uint32_t ao = 2 * j * A_LBO, bo = 2 * j * B_LBO;
    const uint64_t dah = make_smem_desc(ah + ao, A_LBO, SBO), dal = make_smem_desc(al + ao, A_LBO, SBO);
    const uint64_t dbh = make_smem_desc(bh + bo, B_LBO, SBO), dbl = make_smem_desc(bl + bo, B_LBO, SBO);
    umma_tf32(tmem_d, dal, dbh, idesc, (first && j == 0) ? 0u : 1u);   // small cross terms first
    umma_tf32(tmem_d, dah, dbl, idesc, 1u);
    umma_tf32(tmem_d, dah, dbh, idesc, 1u);
  }
}

}  // namespace tc
}  // namespace gpblur
