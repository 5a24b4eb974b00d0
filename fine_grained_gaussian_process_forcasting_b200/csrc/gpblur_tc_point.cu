// Per-point forward / backward of the whitened SVGP on the 5th-gen tensor cores (tcgen05, 3xTF32) for
// padded inducing counts MP in {128, 256}.  One CTA owns a tile of 128 points = the M dimension of the MMA; the
// accumulators live in TMEM (S = x~ z~^T in columns [0, MP), the whitened / back-substituted product in columns
// [MP, 2 MP)); in every epilogue a thread owns ONE point (TMEM lane), so mean / variance / row sums need no
// cross-thread reduction, and the exp()'d cross-covariance goes straight from TMEM registers into the next MMA's
// shared-memory operand planes: it never touches HBM.
//
//   forward  : S = X~ Z~^T -> k = os exp(-1/2 d^2) -> A = k Linv^T (block-triangular: slab s only feeds columns
//              >= 32 s) -> mean, var, sample; A saved for the backward.
//   backward : S again, T = a (diag(c) Linv) (slabs in decreasing order so the first MMA initialises every column)
//              -> kbar = g_mu beta + 2 g_var T, W = kbar o k, r = rowsum(W); W and r go to HBM once.
//   dx       : dx = (W Z~ - r x~) / ell + g_mu w as a third small GEMM (N = Dp) + the per-dimension reductions.
// Operand tiles are produced by the 8 producer warps (coalesced 16-byte loads, TF32 hi/lo split, canonical K-major
// no-swizzle UMMA layout); a 9th warp walks a per-CTA slab table with warp-uniform values and one elected lane issues
// the TMA requests and the MMAs; tcgen05.commit -> mbarrier tracks completion; epilogue chunks are interleaved with
// the slabs (an accumulator chunk is final as soon as the last slab that touches it has retired).
#include <cstdlib>

#include "gpblur_tc.cuh"

namespace gpblur {

namespace {

constexpr int KT = 32;
constexpr int TNP = 128;   // points per tile (MMA M)

struct TcPointArgs {
  WsLayout L;
  void* ws;
  const float* x;
  float* mean;
  float* var;
  float* sample;
  const float* g_mean;
  const float* g_var;
  const float* g_sample;
  const float* var_in;
  float* dx;
  uint64_t seed, offset;
  const unsigned long long* offset_dev;   // optional device-resident addend of `offset` (CUDA-graph replays)
  uint32_t stream_id;
  int ntiles;
  int exp_mode;     // timing experiments only (GPBLUR_TC_EXP): 4 = skip the W chunk stores of the backward (results
                    // are WRONG), bit 3 (8) = no TMA-request polling in the dx issuer (results unchanged)
  long long* trace; // optional event trace buffer (GPBLUR_TRACE_PTR / GPBLUR_FWD_TRACE_PTR = device address)
  long long* dbg;   // optional cycle accounting (GPBLUR_TC_DEBUG=1): [thread 0 | thread 32][16 segments]
};

template <int NB>   // NB = row capacity of the B operand planes
struct Stage {
  static constexpr int A_PLANE = (KT / 4) * TNP * 4;   // floats
  static constexpr int B_PLANE = (KT / 4) * NB * 4;
  static constexpr int FLOATS = 2 * A_PLANE + 2 * B_PLANE;
};

constexpr int kIssuerWarp = kThreads / 32;         // warp 8: the dedicated TMA / MMA issuer
constexpr int kBlockThreads = kThreads + 32;       // 8 producer / epilogue warps + the issuer warp

// named barrier among the 256 producer threads only (the issuer warp never joins it)
__device__ __forceinline__ void prod_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// One pipeline slab = one K = 32 step of one GEMM of the tile: the producers write the A planes of ring stage
// (slab & 1), the TMA engine brings the pre-split B image, the issuer thread fires 3 x 4 MMAs into TMEM.
// The sequence of slabs is the same for every tile, so it is tabulated ONCE per CTA in shared memory and the issuer
// (lane 0 of warp 8) runs a tight loop over the table: its per-slab critical path is only barrier waits, 12
// tcgen05.mma with incrementally updated descriptors, one commit and the next TMA request.
constexpr int kMaxFwdSlabs = 10 * (4 + 8);   // NP (NP + 1) / 2 block pairs x (d-slabs + 8 whitening slabs), M <= 1024
constexpr int kMaxBwdSlabs = 4 * 4 + 32 + 24 + 16 + 8;   // sum over p of (d-slabs + NSL - p SPB)
struct SlabDesc {
  const float* img;    // B image in global memory (hi plane | lo plane), rows x 32 k each
  int rows;            // B rows = MMA N (also sets the leading byte offset of the B descriptor)
  uint32_t tmem_off;   // accumulator column offset inside the CTA's TMEM allocation
  int first;           // bit 0: the first MMA overwrites the accumulator (no accumulate);
                       // bits 8..: 1 + index of the chunk barrier to signal when this slab's MMAs retire (0 = none)
};

// Ring barriers:
//   bars[0..1] mma_done : tcgen05.commit - the MMAs that read stage st have retired (stage reusable)
//   bars[2..3] b_full   : the TMA bulk copy of stage st's B planes has landed (complete_tx)
//   bars[4..5] a_ready  : all 256 producers have written (and proxy-fenced) stage st's A planes
__device__ __forceinline__ void init_ring_barriers(uint64_t* br) {
  tc::mbar_init(&br[0], 1); tc::mbar_init(&br[1], 1);
  tc::mbar_init(&br[2], 1); tc::mbar_init(&br[3], 1);
  tc::mbar_init(&br[4], kThreads); tc::mbar_init(&br[5], kThreads);
  tc::fence_barrier_init();
}

template <int NB>
__device__ __forceinline__ void issue_bulk_b(float* base, uint64_t* bars, int st, const float* image, int rows) {
  float* b_hi = base + st * Stage<NB>::FLOATS + 2 * Stage<NB>::A_PLANE;
  float* b_lo = b_hi + Stage<NB>::B_PLANE;
  const uint32_t bytes = (uint32_t)rows * 128u;          // 8 k-chunks x rows x 16 B
  tc::mbar_expect_tx(&bars[2 + st], 2 * bytes);
  tc::bulk_g2s(b_hi, image, bytes, &bars[2 + st]);
  tc::bulk_g2s(b_lo, image + (size_t)rows * 32, bytes, &bars[2 + st]);
}

// The issuer: the whole warp 8 walks the table with WARP-UNIFORM values (lane-0 broadcasts of the table entries), one
// elected lane executes the tcgen05 / TMA / commit instructions.  Uniform operands live in uniform registers: a
// single thread with per-thread registers costs ~100 cycles per MMA in register -> uniform-register conversion loops
// (measured with scripts/dx_trace.py), more than the execution time of an N <= 128 MMA.
// `ntiles_mine` tiles, each `nslabs` table entries.
// POLL: see the comment at the MMA loop.  Measured on c5 M=256 (same box A/B): dx kernel 0.183 -> 0.171 ms, backward
// kernel 0.272 -> 0.296 ms, forward neutral - so only the dx kernel polls.  The test is additionally predicated on the
// run-time flag `poll_rt` (GPBLUR_TC_EXP bit 3 clears it): with a compile-time-only condition ptxas lays the poll out
// differently and the dx kernel is SLOWER than without polling (0.186 ms).
template <int NB, bool POLL = false>
__device__ __noinline__ void issuer_loop(float* base, uint64_t* bars, uint32_t tmem_base, const SlabDesc* tab, int nslabs,
                                         int ntiles_mine, long long* trace = nullptr, uint64_t* chunk_bars = nullptr, bool poll_rt = true) {
  if (ntiles_mine <= 0 || nslabs <= 0) return;
  constexpr uint64_t kDescHi = ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);      // SBO = 128 B, version 1
  constexpr uint64_t kALbo = (uint64_t)((TNP * 16) >> 4) << 16;                         // A planes: 128 rows
  const uint32_t base_addr = tc::uniform_u32(tc::smem_u32(base));
  tmem_base = tc::uniform_u32(tmem_base);
  nslabs = (int)tc::uniform_u32((uint32_t)nslabs);
  ntiles_mine = (int)tc::uniform_u32((uint32_t)ntiles_mine);
  uint64_t a_hi_desc[2], a_lo_desc[2], b_hi_base[2], b_lo_base[2];
#pragma unroll
  for (int st = 0; st < 2; ++st) {
    const uint32_t a_hi = base_addr + st * Stage<NB>::FLOATS * 4;
    const uint32_t a_lo = a_hi + Stage<NB>::A_PLANE * 4;
    const uint32_t b_hi = a_lo + Stage<NB>::A_PLANE * 4;
    const uint32_t b_lo = b_hi + Stage<NB>::B_PLANE * 4;
    a_hi_desc[st] = kDescHi | kALbo | (uint64_t)(a_hi >> 4);
    a_lo_desc[st] = kDescHi | kALbo | (uint64_t)(a_lo >> 4);
    b_hi_base[st] = kDescHi | (uint64_t)(b_hi >> 4);
    b_lo_base[st] = kDescHi | (uint64_t)(b_lo >> 4);
  }
  auto request_b = [&](int st, int entry) {
    const float* img = reinterpret_cast<const float*>(tc::uniform_u64(reinterpret_cast<uint64_t>(tab[entry].img)));
    const int rows = (int)tc::uniform_u32((uint32_t)tab[entry].rows);
    if (tc::elect_one()) issue_bulk_b<NB>(base, bars, st, img, rows);
    __syncwarp();
  };
  uint32_t uses[2] = {0, 0};
  request_b(0, 0);
  int slab = 0;
  for (int t = 0; t < ntiles_mine; ++t) {
    for (int i = 0; i < nslabs; ++i, ++slab) {
      const int st = slab & 1;
      const uint32_t phase = uses[st] & 1;
      const uint32_t rows = tc::uniform_u32((uint32_t)tab[i].rows);
      const uint32_t tmem_d = tmem_base + tc::uniform_u32(tab[i].tmem_off);
      const uint32_t first_sig = tc::uniform_u32((uint32_t)tab[i].first);
      const uint32_t first = first_sig & 1u, sig = first_sig >> 8;
      const uint32_t idesc = tc::make_idesc_tf32(TNP, (int)rows);
      const uint64_t dbh0 = b_hi_base[st] | ((uint64_t)rows << 16);        // LBO = rows * 16 B
      const uint64_t dbl0 = b_lo_base[st] | ((uint64_t)rows << 16);
      const uint64_t dah0 = a_hi_desc[st], dal0 = a_lo_desc[st];
      const uint64_t bstep = (uint64_t)(2 * rows);                         // two k-chunks of rows * 16 B, >> 4
      long long* tr = (trace && slab < 48 && (threadIdx.x & 31) == 0) ? trace + slab * 6 : nullptr;
      if (tr) tr[0] = clock64();
      tc::mbar_wait(&bars[4 + st], phase);        // A planes written
      if (tr) tr[1] = clock64();
      tc::mbar_wait(&bars[2 + st], phase);        // B image landed
      if (tr) tr[2] = clock64();
      tc::tc_fence_after();
      // TMA request for the next slab's B image into the other stage.  If that stage's last readers have already
      // retired (the usual case: the producers are slower than the tensor core), the request goes out BEFORE this
      // slab's MMAs are issued, so the L2 round trip overlaps the issue time instead of following it.
      const bool more = (i + 1 < nslabs) || (t + 1 < ntiles_mine);
      const int nst = st ^ 1;
      const int nxt = (i + 1 < nslabs) ? i + 1 : 0;
      bool requested = !more;
      if (more) {
        const uint32_t freed = uses[nst] == 0 ? 1u : tc::uniform_u32(tc::mbar_test(&bars[nst], (uses[nst] - 1) & 1) ? 1u : 0u);
        if (freed) {
          request_b(nst, nxt);
          requested = true;
        }
      }
      // The MMA issue BLOCKS once the tensor-core queue is full (12 MMAs take 1400 - 2000 cycles to issue), and the
      // other stage is released by the previous slab's MMAs somewhere in the middle of that.  With POLL the elected
      // thread tests that barrier between k-steps and sends the next slab's TMA request from there instead of after
      // the last MMA (scripts/fwd_trace.py shows the request path).
      if constexpr (POLL) {
        uint32_t req_flag = requested ? 1u : 0u;
        if (tc::elect_one()) {
          const float* nimg = tab[nxt].img;
          const int nrows = tab[nxt].rows;
          const uint32_t nparity = (uses[nst] - 1) & 1;
#pragma unroll
          for (int j = 0; j < KT / 8; ++j) {
            const uint64_t dah = dah0 + (uint64_t)(j * 2 * TNP), dal = dal0 + (uint64_t)(j * 2 * TNP);
            const uint64_t dbh = dbh0 + (uint64_t)j * bstep, dbl = dbl0 + (uint64_t)j * bstep;
            tc::umma_tf32(tmem_d, dal, dbh, idesc, (first && j == 0) ? 0u : 1u);   // small cross terms first
            tc::umma_tf32(tmem_d, dah, dbl, idesc, 1u);
            tc::umma_tf32(tmem_d, dah, dbh, idesc, 1u);
            if (poll_rt && !req_flag && tc::mbar_test(&bars[nst], nparity)) {
              issue_bulk_b<NB>(base, bars, nst, nimg, nrows);
              req_flag = 1u;
            }
          }
          tc::umma_commit(&bars[st]);
          if (sig) tc::umma_commit(&chunk_bars[sig - 1]);
        }
        __syncwarp();
        requested = __any_sync(0xffffffffu, req_flag != 0);
      } else {
        if (tc::elect_one()) {
#pragma unroll
          for (int j = 0; j < KT / 8; ++j) {
            const uint64_t dah = dah0 + (uint64_t)(j * 2 * TNP), dal = dal0 + (uint64_t)(j * 2 * TNP);
            const uint64_t dbh = dbh0 + (uint64_t)j * bstep, dbl = dbl0 + (uint64_t)j * bstep;
            tc::umma_tf32(tmem_d, dal, dbh, idesc, (first && j == 0) ? 0u : 1u);   // small cross terms first
            tc::umma_tf32(tmem_d, dah, dbl, idesc, 1u);
            tc::umma_tf32(tmem_d, dah, dbh, idesc, 1u);
          }
          tc::umma_commit(&bars[st]);
          if (sig) tc::umma_commit(&chunk_bars[sig - 1]);   // an accumulator chunk is final: tell the row-owner warps
        }
        __syncwarp();
      }
      if (tr) tr[3] = clock64();
      uses[st] += 1;
      if (tr) tr[5] = requested ? 0 : 1;          // 1: the request had to wait for the previous slab's MMAs
      if (!requested) {                           // the other stage was still being read: wait, then request
        if (uses[nst] > 0) tc::mbar_wait(&bars[nst], (uses[nst] - 1) & 1);
        request_b(nst, nxt);
      }
      if (tr) tr[4] = clock64();
    }
  }
}

// Producer-side view of the ring (warps 0..7; every producer thread keeps the same counters).
template <int NB>
struct Pipe {
  float* base;
  uint64_t* bars;
  uint32_t uses[2];
  int slab;
  __device__ __forceinline__ void init(float* b, uint64_t* br) { base = b; bars = br; uses[0] = uses[1] = 0; slab = 0; }
  // wait until the MMAs that last read this stage have retired, return its A planes
  __device__ __forceinline__ void acquire(float*& a_hi, float*& a_lo) {
    const int st = slab & 1;
    if (uses[st] > 0) tc::mbar_wait(&bars[st], (uses[st] - 1) & 1);
    a_hi = base + st * Stage<NB>::FLOATS;
    a_lo = a_hi + Stage<NB>::A_PLANE;
  }
  // advance over slabs produced by ANOTHER warp group (both groups count every slab of the shared ring)
  // (the caller must be synchronised with the retirement of the skipped slabs some other way - the row owners wait on
  // the chunk barriers - or the parity waits of acquire() alias)
  __device__ __forceinline__ void skip(int n) {
    for (int i = 0; i < n; ++i) { uses[slab & 1] += 1; slab += 1; }
  }
  // same, but staying within two slabs of the retired MMAs like a producing group does: the parity of a ring barrier
  // is only meaningful one phase ahead, and the slab that follows the skipped ones must not be published before the
  // other group has published (and the tensor core retired) the skipped ones
  __device__ __forceinline__ void skip_wait(int n) {
    for (int i = 0; i < n; ++i) {
      const int st = slab & 1;
      if (uses[st] > 0) tc::mbar_wait(&bars[st], (uses[st] - 1) & 1);
      uses[st] += 1;
      slab += 1;
    }
  }
  // publish this stage's A planes to the issuer (no CTA barrier)
  __device__ __forceinline__ void commit() {
    const int st = slab & 1;
    tc::tc_fence_before();
    tc::fence_async_smem();
    mbar_arrive(&bars[4 + st]);
    uses[st] += 1;
    slab += 1;
  }
  // block until every MMA of the slabs committed so far has completed
  __device__ __forceinline__ void drain() {
    if (slab == 0) return;
    const int st = (slab - 1) & 1;
    tc::mbar_wait(&bars[st], (uses[st] - 1) & 1);
    tc::tc_fence_after();
  }
};

// A [rows x 32 k] K-major operand slab is produced in two steps so that every global load of a slab is in flight
// before the first one is consumed: load_kmajor() fills registers (load4(row, chunk) returns the 4 values
// k = 4 chunk .. 4 chunk + 3 of `row`), store_kmajor() splits them into the TF32 hi / lo planes.  A warp covers
// 8 rows x 4 chunks per pass (64 contiguous bytes per row); stores are conflict-free (8 consecutive rows per
// quarter warp).  `nrows` is a runtime multiple of 8, PLANE_ROWS the plane capacity.
template <int PLANE_ROWS>
struct OpRegs {
  static constexpr int PASSES = (PLANE_ROWS / 8 * 2 + 7) / 8;
  float4 v[PASSES];
};

template <int PLANE_ROWS, class F>
__device__ __forceinline__ void load_kmajor(OpRegs<PLANE_ROWS>& regs, int nrows, F&& load4) {
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 7;   // warp within its 8-warp producer group
  const int rr = lane & 7, cq = lane >> 3;
  const int ntiles = (nrows >> 3) * 2;   // warp tiles: (8-row block, 4-chunk group)
#pragma unroll
  for (int p = 0; p < OpRegs<PLANE_ROWS>::PASSES; ++p) {
    const int wt = warp + 8 * p;
    if (wt < ntiles) regs.v[p] = load4((wt >> 1) * 8 + rr, (wt & 1) * 4 + cq);
  }
}

template <int PLANE_ROWS>
__device__ __forceinline__ void store_kmajor(float* hi, float* lo, const OpRegs<PLANE_ROWS>& regs, int nrows) {
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 7;   // warp within its 8-warp producer group
  const int rr = lane & 7, cq = lane >> 3;
  const int ntiles = (nrows >> 3) * 2;
#pragma unroll
  for (int p = 0; p < OpRegs<PLANE_ROWS>::PASSES; ++p) {
    const int wt = warp + 8 * p;
    if (wt < ntiles) tc::store_split(hi, lo, tc::op_off<PLANE_ROWS>((wt >> 1) * 8 + rr, (wt & 1) * 4 + cq), regs.v[p]);
  }
}

__device__ __forceinline__ float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// x tile row (point), k-chunk `dchunk`: raw() only issues the global load (so that a prefetch one tile ahead does
// not stall on its own data), transform() centres / scales it when the value is consumed.
struct XLoader {
  const float* x; long long n0, N; int D; const float* center; const float* inv_ell; bool vec;
  __device__ __forceinline__ float4 raw(int row, int dchunk) const {
    const long long gn = n0 + row;
    const int d = dchunk * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gn < N && d < D) {
      const float* r = x + (size_t)gn * D;
      if (vec) v = ldg4(r + d);
      else {
        v.x = r[d];
        if (d + 1 < D) v.y = r[d + 1];
        if (d + 2 < D) v.z = r[d + 2];
        if (d + 3 < D) v.w = r[d + 3];
      }
    }
    return v;
  }
  __device__ __forceinline__ float4 transform(float4 v, int row, int dchunk) const {
    const long long gn = n0 + row;
    const int d = dchunk * 4;
    if (gn < N && d < D) {
      const float4 c = ldg4(center + d), ie = ldg4(inv_ell + d);   // padded entries: centre 0, 1 / ell 0
      v.x = (v.x - c.x) * ie.x; v.y = (v.y - c.y) * ie.y; v.z = (v.z - c.z) * ie.z; v.w = (v.w - c.w) * ie.w;
    }
    return v;
  }
};

__device__ __forceinline__ XLoader make_xloader(const TcPointArgs& a, long long n0) {
  const WsLayout& L = a.L;
  return XLoader{a.x, n0, L.N, L.D, ws_cptr<float>(a.ws, L.center), ws_cptr<float>(a.ws, L.inv_ell),
                 (L.D % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0)};
}

// RAW registers of one 32-wide d-slab of the x tile (loads only)
__device__ __forceinline__ void load_x_slab(OpRegs<TNP>& ra, const XLoader& xl, int ds, int DP) {
  load_kmajor<TNP>(ra, TNP, [&](int row, int c) {
    const int dchunk = ds * (KT / 4) + c;
    return dchunk * 4 < DP ? xl.raw(row, dchunk) : make_float4(0.f, 0.f, 0.f, 0.f);
  });
}
// centre / scale the registers of load_x_slab in place (same (row, chunk) mapping as load_kmajor)
__device__ __forceinline__ void transform_x_slab(OpRegs<TNP>& ra, const XLoader& xl, int ds, int DP) {
  const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 7;   // warp within its 8-warp producer group
  const int rr = lane & 7, cq = lane >> 3;
#pragma unroll
  for (int p = 0; p < OpRegs<TNP>::PASSES; ++p) {
    const int wt = warp + 8 * p;
    const int row = (wt >> 1) * 8 + rr, c = (wt & 1) * 4 + cq;
    const int dchunk = ds * (KT / 4) + c;
    if (dchunk * 4 < DP) ra.v[p] = xl.transform(ra.v[p], row, dchunk);
  }
}

// Producer side of S[128, BW] = X~ Z~[block q]^T (TMEM columns [0, BW)): nds slabs of the x tile.  xr0 / xr1: the
// first two d-slabs of this tile's x when `preloaded` (loaded one tile ahead so that the HBM latency hides behind the
// previous tile's MMAs and epilogue).  With `stats`, per-row |x~|^2 and x~ . (ell w) are folded from per-slab
// partials in fixed order (deterministic).
template <int BW>
__device__ __forceinline__ void phase_a(Pipe<BW>& pipe, const TcPointArgs& a, const XLoader& xl, bool stats,
                                        float* part_n, float* part_w, float* xn_s, float* xw_s, bool preloaded,
                                        const OpRegs<TNP>& xr0, const OpRegs<TNP>& xr1, uint64_t* gate, uint32_t blk) {
  const WsLayout& L = a.L;
  const float* wl = ws_cptr<float>(a.ws, L.wl);
  const int DP = L.DP;
  const int nds = DP >= KT ? DP / KT : 1;
  for (int ds = 0; ds < nds; ++ds) {
    OpRegs<TNP> ra;
    if (preloaded && ds == 0) ra = xr0;
    else if (preloaded && ds == 1) ra = xr1;
    else load_x_slab(ra, xl, ds, DP);
    transform_x_slab(ra, xl, ds, DP);
    if (stats) {
      // row-statistic partials of the chunks this thread just loaded (same (row, chunk) mapping as load_kmajor)
      const int lane = threadIdx.x & 31, warp = (threadIdx.x >> 5) & 7;   // warp within its 8-warp producer group
      const int rr = lane & 7, cq = lane >> 3;
#pragma unroll
      for (int p = 0; p < OpRegs<TNP>::PASSES; ++p) {
        const int wt = warp + 8 * p;
        const int row = (wt >> 1) * 8 + rr, c = (wt & 1) * 4 + cq;
        const int dchunk = ds * (KT / 4) + c;
        const float4 v = ra.v[p];
        const float4 w4 = dchunk * 4 < DP ? ldg4(wl + dchunk * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        part_n[c * TNP + row] = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        part_w[c * TNP + row] = v.x * w4.x + v.y * w4.y + v.z * w4.z + v.w * w4.w;
      }
    }
    float *a_hi, *a_lo;
    pipe.acquire(a_hi, a_lo);
    store_kmajor<TNP>(a_hi, a_lo, ra, TNP);
    // `gate[ds / 2]` (one phase per block): the other producer group has passed this pair of slabs in its own slab
    // count (see the loader warps of the backward kernel).  It gets there long before; the wait only rules out
    // parity aliasing on the ring barriers.
    if ((ds & 1) == 0) tc::mbar_wait(&gate[ds >> 1], blk & 1);
    pipe.commit();
    if (stats) {
      // fold this slab's 8 chunk partials in fixed order (bit-deterministic)
      prod_sync();
      if (threadIdx.x < TNP) {
        float n2 = ds == 0 ? 0.f : xn_s[threadIdx.x], xw = ds == 0 ? 0.f : xw_s[threadIdx.x];
#pragma unroll
        for (int c = 0; c < KT / 4; ++c) {
          n2 += part_n[c * TNP + threadIdx.x];
          xw += part_w[c * TNP + threadIdx.x];
        }
        xn_s[threadIdx.x] = n2;
        xw_s[threadIdx.x] = xw;
      }
      prod_sync();
    }
  }
}

// ---- slab tables (one entry per pipeline slab of a tile, in issue order) ----
// forward: for p, for q <= p: nds slabs of S = X~ Z~[q]^T, then SPB whitening slabs of block q into output block p
template <int BW>
__device__ __forceinline__ int fwd_table(SlabDesc* tab, const TcPointArgs& a) {
  const WsLayout& L = a.L;
  const int MP = L.MP, NP = MP / BW, SPB = BW / KT;
  const int nds = L.DP >= KT ? L.DP / KT : 1;
  const float* ZtU = ws_cptr<float>(a.ws, L.ZtU);
  const float* LinvU = ws_cptr<float>(a.ws, L.LinvU);
  const int per = nds + SPB, total = NP * (NP + 1) / 2 * per;
  for (int i = threadIdx.x; i < total; i += kThreads) {
    const int pair = i / per, j = i - pair * per;
    int p = 0;
    while ((p + 1) * (p + 2) / 2 <= pair) ++p;
    const int q = pair - p * (p + 1) / 2;
    SlabDesc d;
    if (j < nds) {
      d.img = ZtU + tc_zt_image(MP, nds, q, j);
      d.rows = BW; d.tmem_off = 0; d.first = j == 0;
    } else {
      const int sl = j - nds, sg = q * SPB + sl;
      int rows;
      d.img = LinvU + tc_linv_image(MP, p, sg, &rows);
      d.rows = rows; d.tmem_off = (uint32_t)(BW + BW - rows);
      d.first = (q == 0 && sl == 0) ? 1 : 0;
    }
    tab[i] = d;
  }
  return total;
}
// backward: for p: nds slabs of S (block p), then T[:, block p] slabs s = NSL - 1 ... p SPB (decreasing)
template <int BW>
__device__ __forceinline__ int bwd_table(SlabDesc* tab, const TcPointArgs& a) {
  const WsLayout& L = a.L;
  const int MP = L.MP, NP = MP / BW, SPB = BW / KT, NSL = MP / KT;
  const int nds = L.DP >= KT ? L.DP / KT : 1;
  const float* ZtU = ws_cptr<float>(a.ws, L.ZtU);
  const float* LCTU = ws_cptr<float>(a.ws, L.LCTU);
  int total = 0;
  for (int p = 0; p < NP; ++p) total += nds + NSL - p * SPB;
  for (int i = threadIdx.x; i < total; i += kThreads) {
    int p = 0, base = 0;
    while (i >= base + nds + NSL - p * SPB) { base += nds + NSL - p * SPB; ++p; }
    const int j = i - base;
    SlabDesc d;
    if (j < nds) {
      d.img = ZtU + tc_zt_image(MP, nds, p, j);
      d.rows = BW; d.tmem_off = 0; d.first = j == 0;
    } else {
      const int sg = NSL - 1 - (j - nds);
      int rows;
      d.img = LCTU + tc_lct_image(MP, p, sg, &rows);
      const int c = sg - p * SPB;                   // chunk finalised by this slab (slabs run in decreasing order)
      d.rows = rows; d.tmem_off = (uint32_t)BW; d.first = (sg == NSL - 1 ? 1 : 0) | (c < SPB ? ((c + 1) << 8) : 0);
    }
    tab[i] = d;
  }
  return total;
}

// Cross-covariance values from the S accumulators (inlined where used):
//   k = os exp(-1/2 max(|x|^2 + |z|^2 - 2 s, 0)) = 2^min(s log2e + xnc + znc[m], log2 os)
// with xnc = -1/2 log2e |x~|^2 and znc[m] = -1/2 log2e |z~_m|^2 + log2 os (-1e30 on padded columns => k = 0):
// 4 FP32 instructions + one MUFU per element.

// Pitch (floats) of a per-warp [32 rows x 32 columns] shared-memory staging tile: conflict-free 16-byte accesses.
constexpr int kStagePitch = 36;
// =================================================================================================
// forward.  The inducing dimension is processed in column blocks of width BW (= min(MP, 256), the TMEM budget:
// S in columns [0, BW), the whitened product of the current output block in [BW, 2 BW)).  Output block p needs the
// cross-covariance blocks q <= p (Linv is lower triangular); for MP > 256 block q is recomputed for every p >= q.
// =================================================================================================
// Warp roles (544 threads): TWO producer groups (warps 0..7 and 8..15) that own ALTERNATE pipeline slabs - group g
// always writes ring stage g - and the issuer warp 16.  The per-slab work of a producer is a chain of latencies
// (TMEM load -> exp -> stage acquire -> shared-memory stores under UMMA operand traffic -> proxy fence -> arrive) that
// eight warps cannot hide: 2000 cycles per slab against 1536 cycles of MMAs (N = 256).  With two groups each chain
// has two slab times, and the chunk epilogues (chunk c is final once whitening slab c of the pass q == p has retired,
// which the owner of slab c + 2 learns from its stage acquire) are spread over both groups as well.
// Both groups count every slab, but a group only ever waits on the barriers of its OWN stage (its phase can only move
// when the group itself commits) - with one exception, the wait for S after phase A, whose last slab may belong to
// the other group: the 512-thread barrier in front of it guarantees that the owner has seen the stage's previous use
// retire, and the stage cannot advance another phase before this group publishes the slab in between.  (Waiting on
// the other group's stage anywhere else can observe the barrier TWO phases later and alias.)
constexpr int kTwoGroupThreads = 2 * kThreads + 32;
constexpr int kTwoGroupIssuerWarp = 2 * kThreads / 32;
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 256;" ::"r"(id) : "memory"); }
// all 512 producer threads (the issuer warp never joins)
__device__ __forceinline__ void producers_sync() { asm volatile("bar.sync 3, 512;" ::: "memory"); }

// Store of a [32 rows x 16 columns] half chunk held one row per lane (16 consecutive floats), without staging:
// neighbouring lanes exchange two 16-byte pieces, so that every store instruction writes 32 contiguous bytes per lane
// pair - complete sectors - instead of half sectors of 32 different rows.  `dst` = global address of (row 0 of the
// warp, first column), `ld` = row pitch in floats, `nvalid` rows exist.
__device__ __forceinline__ void warp_store_rows16(float* dst, size_t ld, const float (&v)[16], int lane, int nvalid) {
  const bool odd = lane & 1;
  float sx[8], rx[8];
#pragma unroll
  for (int i = 0; i < 4; ++i) {          // even lanes send pieces 1 and 3, odd lanes pieces 0 and 2
    sx[i] = odd ? v[i] : v[4 + i];
    sx[4 + i] = odd ? v[8 + i] : v[12 + i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) rx[i] = __shfl_xor_sync(0xffffffffu, sx[i], 1);
  const int row_a = lane & ~1, row_b = row_a + 1;
  float* pa = dst + (size_t)row_a * ld + (odd ? 4 : 0);
  float* pb = dst + (size_t)row_b * ld + (odd ? 4 : 0);
  if (row_a < nvalid) {
    *reinterpret_cast<float4*>(pa) = odd ? make_float4(rx[0], rx[1], rx[2], rx[3]) : make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(pa + 8) =
        odd ? make_float4(rx[4], rx[5], rx[6], rx[7]) : make_float4(v[8], v[9], v[10], v[11]);
  }
  if (row_b < nvalid) {
    *reinterpret_cast<float4*>(pb) = odd ? make_float4(v[4], v[5], v[6], v[7]) : make_float4(rx[0], rx[1], rx[2], rx[3]);
    *reinterpret_cast<float4*>(pb + 8) =
        odd ? make_float4(v[12], v[13], v[14], v[15]) : make_float4(rx[4], rx[5], rx[6], rx[7]);
  }
}

template <int BW>
__global__ void __launch_bounds__(kTwoGroupThreads, 1) tc_point_fwd_kernel(TcPointArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bars[6];
  __shared__ uint32_t tmem_slot;
  __shared__ SlabDesc tab[kMaxFwdSlabs];
  __shared__ int tab_n;
  __shared__ float zn_s[BW], m_s[BW], c_s[BW];      // exponent offsets of block q; m, c = s^2 - 1 of block p
  constexpr int kMaxDs = 4;                         // d-slabs of the x tile (D <= 128)
  __shared__ float part_n[kMaxDs][2][TNP], part_w[kMaxDs][2][TNP];   // row-statistic partials per d-slab / chunk half
  __shared__ float mu_s[4][TNP], vv_s[4][TNP];      // mean / variance partials per (group, column half)

  const WsLayout& L = a.L;
  const int MP = L.MP;
  const int NP = MP / BW;
  constexpr int SPB = BW / KT;
  const long long N = L.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* hyp = ws_cptr<float>(a.ws, L.hyp);
  float* Ag = ws_ptr<float>(a.ws, L.A);
  const float* znc_g = ws_cptr<float>(a.ws, L.znc);
  const float* mvec_g = ws_cptr<float>(a.ws, L.mvec);
  const float* cvec_g = ws_cptr<float>(a.ws, L.cvec);
  const float* wl = ws_cptr<float>(a.ws, L.wl);
  const float os = hyp[H_OS], jit = hyp[H_JIT], cwb = hyp[H_CWB];
  const float l2os = log2f(os);
  const int DP = L.DP;
  const int nds = DP >= KT ? DP / KT : 1;

  constexpr uint32_t TMEM_COLS = 2 * BW;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) init_ring_barriers(bars);
  if (tid < kThreads) {
    for (int i = tid; i < BW; i += kThreads) {      // block 0; reloaded per (p, q) when MP > BW
      zn_s[i] = znc_g[i];                            // exponent offsets (formula above)
      m_s[i] = mvec_g[i];
      c_s[i] = cvec_g[i];
    }
    const int n = fwd_table<BW>(tab, a);
    if (tid == 0) tab_n = n;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_s = tmem_slot, tmem_a = tmem_slot + BW;
  const int tiles_mine = a.ntiles > (int)blockIdx.x ? (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp == kTwoGroupIssuerWarp) {
    // ---------------- issuer warp: one elected thread drives the TMA requests and the tensor core ----------------
    issuer_loop<BW>(stage_base, bars, tmem_slot, tab, tab_n, tiles_mine, blockIdx.x == 0 ? a.trace : nullptr);
  } else {
    // ---------------- producer groups ----------------
    const int g = warp >> 3;                          // group = ring stage this thread writes
    const int quad = warp & 3, half = (warp >> 2) & 1;   // TMEM lane quadrant / column half of this warp
    const int row = quad * 32 + lane;                 // the point this thread owns
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    const int slabs_per_tile = tab_n;
    Pipe<BW> pipe;
    pipe.init(stage_base, bars);
    long long seg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long tlast = clock64();
#define SEG(i) do { if (a.dbg) { const long long tnow = clock64(); seg[i] += tnow - tlast; tlast = tnow; } } while (0)
    // the first d-slab of a tile that this group owns (slab parity) is loaded one tile ahead
    OpRegs<TNP> xr;
    int ds_pre = g;                                   // tile 0 starts at slab 0: group g owns d-slab g
    if (ds_pre < nds) load_x_slab(xr, make_xloader(a, (long long)blockIdx.x * TNP), ds_pre, DP);
    // ---- phase A of one pass: S[128, BW] = X~ Z~[block q]^T, the d-slabs alternate between the groups.  (Issuing it
    // one pass ahead, before the tail epilogue of the previous block, was measured: no gain - the tensor pipe, not
    // the producers, paces these kernels.) ----
    auto phase_a_slabs = [&](const XLoader& xl, bool first_pass) {
      for (int ds = 0; ds < nds; ++ds) {
        if ((pipe.slab & 1) != g) { pipe.skip(1); continue; }
        OpRegs<TNP> ra;
        if (first_pass && ds == ds_pre) ra = xr;
        else load_x_slab(ra, xl, ds, DP);
        transform_x_slab(ra, xl, ds, DP);
        if (first_pass) {
          // row-statistic partials (same (row, chunk) mapping as load_kmajor): the four chunk groups of a row sit
          // in lanes rr, rr + 8, rr + 16, rr + 24 (fixed shuffle tree), the two chunk halves in neighbouring warps
          const int gw = warp & 7, rr = lane & 7, cq = lane >> 3;
#pragma unroll
          for (int ps = 0; ps < OpRegs<TNP>::PASSES; ++ps) {
            const int wt = gw + 8 * ps;
            const int prow = (wt >> 1) * 8 + rr, c = (wt & 1) * 4 + cq;
            const int dchunk = ds * (KT / 4) + c;
            const float4 v = ra.v[ps];
            const float4 w4 = dchunk * 4 < DP ? ldg4(wl + dchunk * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            // explicit FMA chains: the lambda is inlined at two call sites, and the outputs must not depend on how
            // the compiler contracts each copy (shard invariance is tested bit-exactly)
            float pn = fmaf(v.w, v.w, fmaf(v.z, v.z, fmaf(v.y, v.y, v.x * v.x)));
            float pw = fmaf(v.w, w4.w, fmaf(v.z, w4.z, fmaf(v.y, w4.y, v.x * w4.x)));
            pn += __shfl_xor_sync(0xffffffffu, pn, 8);
            pw += __shfl_xor_sync(0xffffffffu, pw, 8);
            pn += __shfl_xor_sync(0xffffffffu, pn, 16);
            pw += __shfl_xor_sync(0xffffffffu, pw, 16);
            if (cq == 0) { part_n[ds][wt & 1][prow] = pn; part_w[ds][wt & 1][prow] = pw; }
          }
        }
        float *a_hi, *a_lo;
        pipe.acquire(a_hi, a_lo);
        store_kmajor<TNP>(a_hi, a_lo, ra, TNP);
        pipe.commit();
      }
    };
    bool have_prev = false;                           // deferred mean / variance / sample of the previous tile
    long long prev_gn = 0;
    float xw_prev = 0.f;
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      const long long n0 = (long long)tile * TNP;
      const XLoader xl = make_xloader(a, n0);
      const bool more_tiles = tile + (int)gridDim.x < a.ntiles;
      const long long w0 = n0 + quad * 32;        // first point of this warp
      const int nvalid = N - w0 >= 32 ? 32 : (N > w0 ? (int)(N - w0) : 0);
      float mu = 0.f, vv = 0.f, xnc = 0.f;
      SEG(7);
      // epilogue of one 32-column chunk c of output block p (16 columns per column half): mean / variance partials
      // of the own point, A saved for the backward
      auto epi_chunk = [&](int p, int c) {
        const int col = c * 32 + half * 16;
        float v[16];
        tc::tmem_ld16(tmem_a + lane_base + (uint32_t)col, v);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          mu = fmaf(v[i], m_s[col + i], mu);
          vv = fmaf(c_s[col + i] * v[i], v[i], vv);
        }
        if (L.training) warp_store_rows16(Ag + (size_t)w0 * MP + p * BW + col, MP, v, lane, nvalid);
      };
      for (int p = 0; p < NP; ++p) {
        for (int q = 0; q <= p; ++q) {
          const bool first_pass = p == 0 && q == 0;
          if (NP > 1) {                               // per-block constants (single block: loaded once at kernel start)
            tc::tc_fence_before();
            producers_sync();
            if (tid < BW) {
              zn_s[tid] = znc_g[q * BW + tid];
              if (q == 0) { m_s[tid] = mvec_g[p * BW + tid]; c_s[tid] = cvec_g[p * BW + tid]; }
            }
          }
          phase_a_slabs(xl, first_pass);
          // every producer has (a) read the previous output block out of the A accumulators - the first whitening
          // slab below overwrites them -, (b) published its row-statistic partials / the block constants
          tc::tc_fence_before();
          producers_sync();
          tc::tc_fence_after();
          if (first_pass) {
            float n2 = 0.f, xw = 0.f;
            for (int ds = 0; ds < nds; ++ds) {        // fixed order (bit-deterministic)
              n2 += part_n[ds][0][row] + part_n[ds][1][row];
              xw += part_w[ds][0][row] + part_w[ds][1][row];
            }
            xnc = -0.72134752044448170f * n2;
            if (g == 0 && half == 0) {
              if (have_prev && prev_gn < N) {         // the previous tile's partials are complete (barrier above)
                const float mean = mu_s[0][row] + mu_s[1][row] + mu_s[2][row] + mu_s[3][row] + xw_prev + cwb;
                const float var = fmaxf(os + jit + vv_s[0][row] + vv_s[1][row] + vv_s[2][row] + vv_s[3][row], kMinVariance);
                a.mean[prev_gn] = mean;
                a.var[prev_gn] = var;
                if (a.sample)
                  a.sample[prev_gn] = fmaf(sqrtf(var), philox_normal(a.seed, rng_offset(a.offset, a.offset_dev) + (uint64_t)prev_gn, a.stream_id), mean);
              }
              xw_prev = xw;
            }
            if (more_tiles) {                         // next tile's first own d-slab: in flight during this tile
              ds_pre = (((pipe.slab - nds + slabs_per_tile) & 1) == g) ? 0 : 1;
              if (ds_pre < nds) load_x_slab(xr, make_xloader(a, n0 + (long long)gridDim.x * TNP), ds_pre, DP);
            }
          }
          SEG(0);                                     // phase A (x split, A planes published, barrier)
          pipe.drain();                               // S of block q complete
          SEG(1);
          // ---- whitening: A[:, block p] += k[:, slab s] Linv[block p rows >= 32 s, slab s]^T ----
          // this group owns the slabs sl = f, f + 2, ...; the S columns of its next slab are requested from TMEM
          // before the current one is exponentiated (two register sets)
          const int f = ((pipe.slab & 1) == g) ? 0 : 1;
          uint32_t sreg[2][16];
          tc::tmem_ld16_issue(tmem_s + lane_base + (uint32_t)(f * KT + half * 16), sreg[0]);
#pragma unroll
          for (int i = 0; i < SPB / 2; ++i) {
            if (f == 1) pipe.skip(1);
            const int sl = 2 * i + f;
            // fused epilogue of S: the two column halves of a lane quadrant take 16 columns each
            float v[16];
            const int col0 = sl * KT + half * 16;
            tc::tmem_ld16_wait(sreg[i & 1]);
            if (i + 1 < SPB / 2) tc::tmem_ld16_issue(tmem_s + lane_base + (uint32_t)(col0 + 2 * KT), sreg[(i + 1) & 1]);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              v[j] = tc::ex2_approx(fminf(fmaf(__uint_as_float(sreg[i & 1][j]), 1.4426950408889634f,
                                               xnc + zn_s[col0 + j]), l2os));
            SEG(2);                                   // TMEM load + exp
            long long* ptr = (a.trace && blockIdx.x == 0 && (tid & 255) == 0 && pipe.slab < 48) ? a.trace + 512 + pipe.slab * 4 : nullptr;
            if (ptr) ptr[0] = clock64();
            float *a_hi, *a_lo;
            pipe.acquire(a_hi, a_lo);                 // slab sl - 2 (this group's previous one) has retired
            if (ptr) ptr[1] = clock64();
            SEG(3);
#pragma unroll
            for (int c = 0; c < 4; ++c)               // k-chunks half * 4 + c of the slab
              tc::store_split(a_hi, a_lo, tc::op_off<TNP>(row, half * 4 + c),
                              make_float4(v[c * 4 + 0], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]));
            SEG(4);                                   // split + store of the k slab
            if (ptr) ptr[2] = clock64();
            pipe.commit();
            if (ptr) ptr[3] = clock64();
            SEG(5);                                   // fences + arrive
            if (q == p && sl >= 2) {                  // chunk sl - 2 is final: handled while the tensor core works on
              tc::tc_fence_after();
              epi_chunk(p, sl - 2);
              SEG(6);
            }
            if (f == 0) pipe.skip(1);
          }
          if (q == p) {
            // this group's last slab of the block: its chunk is final once it has retired
            tc::mbar_wait(&bars[g], (pipe.uses[g] - 1) & 1);
            tc::tc_fence_after();
            epi_chunk(p, SPB - 2 + f);
            SEG(6);
          }
        }
      }
      mu_s[g * 2 + half][row] = mu;
      vv_s[g * 2 + half][row] = vv;
      have_prev = true;
      prev_gn = n0 + row;
    }
    tc::tc_fence_before();
    producers_sync();
    if (g == 0 && half == 0 && have_prev && prev_gn < N) {
      const float mean = mu_s[0][row] + mu_s[1][row] + mu_s[2][row] + mu_s[3][row] + xw_prev + cwb;
      const float var = fmaxf(os + jit + vv_s[0][row] + vv_s[1][row] + vv_s[2][row] + vv_s[3][row], kMinVariance);
      a.mean[prev_gn] = mean;
      a.var[prev_gn] = var;
      if (a.sample)
        a.sample[prev_gn] = fmaf(sqrtf(var), philox_normal(a.seed, rng_offset(a.offset, a.offset_dev) + (uint64_t)prev_gn, a.stream_id), mean);
    }
    if (a.dbg && blockIdx.x == 0 && (tid == 0 || tid == 32))
      for (int i = 0; i < 8; ++i) a.dbg[(tid == 0 ? 0 : 8) + i] = seg[i];
#undef SEG
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_slot, TMEM_COLS);
}

// =================================================================================================
// backward: W = kbar o k and its row sums, one column block p of width BW at a time
// =================================================================================================
// Warp roles of the backward kernel (640 threads, 96 registers each; warps 17..19 idle):
//   warps 0..7   ROW OWNERS : phase A (x~ planes, row statistics), epilogue chunks (thread = point), W stores
//   warps 8..15  LOADERS    : the saved-A slabs of the T GEMM: global loads -> TF32 split -> A planes -> arrive
//   warp 16      ISSUER     : TMA requests + MMAs (+ per-chunk commits)
// The two producer groups share one operand ring; both count every slab, each arrives (256 threads) only on its own.
constexpr int kBwdThreads = 640;   // (measured: 0.30 ms at 640 threads against 0.33 ms at 544, c5 M=256)
constexpr int kBwdIssuerWarp = kTwoGroupIssuerWarp;

template <int BW>
__global__ void __launch_bounds__(kBwdThreads, 1) tc_point_bwd_kernel(TcPointArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bars[6];
  __shared__ __align__(8) uint64_t chunk_bars[8];                    // accumulator chunk c of the current block is final
  __shared__ __align__(8) uint64_t ld_ready[2];                      // the loader warps have passed phase-A slabs 0-1 / 2-3 of the block
  __shared__ uint32_t tmem_slot;
  __shared__ SlabDesc tab[kMaxBwdSlabs];
  __shared__ int tab_n;
  __shared__ float zn_s[BW], beta_s[BW];                              // exponent offsets / beta of column block p
  __shared__ float xn_s[TNP], xw_s[TNP], r_s[TNP];
  __shared__ float part_n[(KT / 4) * TNP], part_w[(KT / 4) * TNP];  // per-slab row-statistic partials of phase A

  const WsLayout& L = a.L;
  const int MP = L.MP;
  const int NP = MP / BW, SPB = BW / KT, NSL = MP / KT;
  const long long N = L.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* hyp = ws_cptr<float>(a.ws, L.hyp);
  const float* Ag = ws_cptr<float>(a.ws, L.A);
  float* Wg = ws_ptr<float>(a.ws, L.W);
  float* gsc = ws_ptr<float>(a.ws, L.gsc);
  float* rrow = ws_ptr<float>(a.ws, L.rrow);
  const float* znc_g = ws_cptr<float>(a.ws, L.znc);
  const float* beta_g = ws_cptr<float>(a.ws, L.beta);
  const float l2os = log2f(hyp[H_OS]);
  const int nds = L.DP >= KT ? L.DP / KT : 1;

  constexpr uint32_t TMEM_COLS = 2 * BW;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) {
    init_ring_barriers(bars);
    for (int i = 0; i < 8; ++i) tc::mbar_init(&chunk_bars[i], 1);
    tc::mbar_init(&ld_ready[0], kThreads);
    tc::mbar_init(&ld_ready[1], kThreads);
    tc::fence_barrier_init();
  }
  if (tid < kThreads) {
    for (int i = tid; i < BW; i += kThreads) {      // block 0; reloaded per p when MP > BW
      zn_s[i] = znc_g[i];                            // exponent offsets (formula above)
      beta_s[i] = beta_g[i];
    }
    const int n = bwd_table<BW>(tab, a);
    if (tid == 0) tab_n = n;
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_s = tmem_slot, tmem_t = tmem_slot + BW;
  const int tiles_mine = a.ntiles > (int)blockIdx.x ? (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

  if (warp >= kBwdIssuerWarp) {
    // ---------------- issuer warpgroup ----------------
    if (warp == kBwdIssuerWarp) issuer_loop<BW>(stage_base, bars, tmem_slot, tab, tab_n, tiles_mine, nullptr, chunk_bars);
  } else if (warp >= 8) {
    // ---------------- loader warps: saved-A slabs -> A planes ----------------
    Pipe<BW> pipe;
    pipe.init(stage_base, bars);
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      const long long n0 = (long long)tile * TNP;
      auto load_a = [&](OpRegs<TNP>& regs, int sl) {
        load_kmajor<TNP>(regs, TNP, [&](int r, int c) {
          long long g2 = n0 + r;
          if (g2 >= N) g2 = N - 1;                       // clamped rows carry g = 0
          return ldg4(Ag + (size_t)g2 * MP + sl * KT + c * 4);
        });
      };
      for (int p = 0; p < NP; ++p) {
        const int s_lo = p * SPB;
        // three slabs in flight (rotating register sets); the first ones are requested while the row owners are
        // still in phase A
        OpRegs<TNP> r0, r1, r2;
        load_a(r0, NSL - 1);
        if (NSL - 2 >= s_lo) load_a(r1, NSL - 2);
        if (NSL - 3 >= s_lo) load_a(r2, NSL - 3);
        // the phase-A slabs of this block belong to the row owners.  skip_wait() tests the PREVIOUS use of the slab's
        // stage; the row owners publish a pair of slabs only after the arrival for it, so no loader can find a ring barrier two
        // phases further than it expects (parity aliasing)
        for (int ds = 0; ds < nds; ds += 2) {
          pipe.skip_wait(nds - ds < 2 ? nds - ds : 2);
          mbar_arrive(&ld_ready[ds >> 1]);
        }
        for (int s = NSL - 1; s >= s_lo; s -= 3) {
          float *a_hi, *a_lo;
          pipe.acquire(a_hi, a_lo);
          store_kmajor<TNP>(a_hi, a_lo, r0, TNP);
          if (s - 3 >= s_lo) load_a(r0, s - 3);
          pipe.commit();
          if (s - 1 < s_lo) break;
          pipe.acquire(a_hi, a_lo);
          store_kmajor<TNP>(a_hi, a_lo, r1, TNP);
          if (s - 4 >= s_lo) load_a(r1, s - 4);
          pipe.commit();
          if (s - 2 < s_lo) break;
          pipe.acquire(a_hi, a_lo);
          store_kmajor<TNP>(a_hi, a_lo, r2, TNP);
          if (s - 5 >= s_lo) load_a(r2, s - 5);
          pipe.commit();
        }
      }
    }
  } else {
    // ---------------- row-owner warps ----------------
    Pipe<BW> pipe;
    pipe.init(stage_base, bars);
    const int quad = warp & 3, half = warp >> 2;
    const int row = quad * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
    uint32_t blk = 0;                                 // blocks processed so far: phase of the chunk barriers

    long long bseg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long blast = clock64();
#define BSEG(i) do { if (a.dbg) { const long long tnow = clock64(); bseg[i] += tnow - blast; blast = tnow; } } while (0)
    OpRegs<TNP> xr0, xr1;
    {
      const XLoader xl0 = make_xloader(a, (long long)blockIdx.x * TNP);
      load_x_slab(xr0, xl0, 0, L.DP);
      if (L.DP > KT) load_x_slab(xr1, xl0, 1, L.DP);
    }
    for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
      const long long n0 = (long long)tile * TNP;
      const long long gn = n0 + row;
      const XLoader xl = make_xloader(a, n0);
      const bool more_tiles = tile + (int)gridDim.x < a.ntiles;
      const long long w0 = n0 + quad * 32;        // first point of this warp
      const int nvalid = N - w0 >= 32 ? 32 : (N > w0 ? (int)(N - w0) : 0);
      // ---- fold the upstream gradients of this thread's point ----
      float gm = 0.f, gv = 0.f;
      if (gn < N) {
        if (a.g_mean) gm = a.g_mean[gn];
        if (a.g_var) gv = a.g_var[gn];
        const float v = a.var_in[gn];
        if (a.g_sample) {
          const float gs = a.g_sample[gn];
          const float eps = philox_normal(a.seed, rng_offset(a.offset, a.offset_dev) + (uint64_t)gn, a.stream_id);
          gm += gs;
          gv = fmaf(gs * eps, 0.5f * rsqrtf(v), gv);
        }
        if (v <= kMinVariance) gv = 0.f;
        if (half == 0) { gsc[gn] = gm; gsc[N + gn] = gv; }
      }
      float rsum = 0.f, xnc = 0.f;
      BSEG(0);                                        // tile head (upstream gradients)
      // epilogue of one 32-column chunk c of column block p (16 columns per column half): W = kbar o k, its row sum,
      // W saved for the dx / W^T X kernels.  T[:, chunk c] is last touched by slab p SPB + c, the slabs run in
      // DEcreasing order, and the issuer commits that slab onto chunk_bars[c]: the chunks are handled while the
      // tensor core (and the loader warps) work on the remaining slabs.
      auto epi_chunk = [&](int p, int c) {
        const int col = c * 32 + half * 16;
        uint32_t kr[16], tr[16];
        tc::tmem_ld16_issue(tmem_s + lane_base + (uint32_t)col, kr);
        tc::tmem_ld16_issue(tmem_t + lane_base + (uint32_t)col, tr);
        tc::tmem_ld16_wait(kr);
        tc::tmem_ld16_wait(tr);
        float t[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float k = tc::ex2_approx(fminf(fmaf(__uint_as_float(kr[i]), 1.4426950408889634f, xnc + zn_s[col + i]), l2os));
          const float kb = fmaf(2.0f * gv, __uint_as_float(tr[i]), gm * beta_s[col + i]);
          t[i] = kb * k;
          rsum += t[i];
        }
        if (a.exp_mode != 4)
          warp_store_rows16(Wg + (size_t)w0 * MP + p * BW + col, MP, t, lane, nvalid);   // (no staging: lane-pair exchange)
      };
      for (int p = 0; p < NP; ++p, ++blk) {
        if (NP > 1) {                                 // per-block constants (single block: loaded once at kernel start)
          group_sync(1);
          if (tid < BW) { zn_s[tid] = znc_g[p * BW + tid]; beta_s[tid] = beta_g[p * BW + tid]; }
          group_sync(1);
        }
        phase_a<BW>(pipe, a, xl, p == 0, part_n, part_w, xn_s, xw_s, p == 0, xr0, xr1, ld_ready, blk);
        if (p == 0 && more_tiles) {
          const XLoader xln = make_xloader(a, n0 + (long long)gridDim.x * TNP);
          load_x_slab(xr0, xln, 0, L.DP);
          if (L.DP > KT) load_x_slab(xr1, xln, 1, L.DP);
        }
        xnc = -0.72134752044448170f * xn_s[row];
        BSEG(1);                                      // phase A
        pipe.skip(NSL - p * SPB);                     // the T slabs of this block belong to the loader warps
        for (int c = SPB - 1; c >= 0; --c) {
          tc::mbar_wait(&chunk_bars[c], blk & 1);
          tc::tc_fence_after();
          BSEG(2);                                    // wait for the chunk's last slab
          epi_chunk(p, c);
          BSEG(5);                                    // epilogue chunk
        }
        // the TMEM reads are done before ANY row owner publishes A planes of the next phase (its MMAs overwrite S)
        tc::tc_fence_before();
        group_sync(1);
      }
      if (half == 1) r_s[row] = rsum;
      group_sync(1);
      if (half == 0 && gn < N) rrow[gn] = rsum + r_s[row];
      group_sync(1);
      BSEG(7);                                        // row sums
    }
    if (a.dbg && blockIdx.x == 0 && tid == 0)
      for (int i = 0; i < 8; ++i) a.dbg[16 + i] = bseg[i];
#undef BSEG
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_slot, TMEM_COLS);
}

// =================================================================================================
// dx = (W Z~ - r x~) / ell + g_mu w  and the per-dimension reductions q, wbar + scalar sums
// =================================================================================================
template <int DPT>   // DPT = MMA N = padded input dim (32, 64 or 128)
__global__ void __launch_bounds__(kBlockThreads, 1) tc_dx_kernel(TcPointArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bars[6];
  __shared__ uint32_t tmem_slot;
  __shared__ SlabDesc tab[GPBLUR_MAX_M / KT];
  __shared__ __align__(16) float red[8][32][kStagePitch];   // per-warp staging tile (x~ in, dx out)
  __shared__ float rg_s[8][2][32];                             // per-warp r[n], g_mu[n] of its 32 points
  __shared__ float part_q[8][32], part_t[8][32], part_q2[8][32], part_t2[8][32], part_sc[8][4];
  __shared__ float q_s[DPT], t1_s[DPT], sc_s[4];

  const WsLayout& L = a.L;
  const int D = L.D, DP = L.DP, MP = L.MP;
  const long long N = L.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* inv_ell = ws_cptr<float>(a.ws, L.inv_ell);
  const float* ellv = ws_cptr<float>(a.ws, L.ell);
  const float* center = ws_cptr<float>(a.ws, L.center);
  const float* wl = ws_cptr<float>(a.ws, L.wl);
  const float* ZtTU = ws_cptr<float>(a.ws, L.ZtTU);
  const float* Wg = ws_cptr<float>(a.ws, L.W);
  const float* gsc = ws_cptr<float>(a.ws, L.gsc);
  const float* rrow = ws_cptr<float>(a.ws, L.rrow);
  float* vecpart = ws_ptr<float>(a.ws, L.vecpart);

  constexpr uint32_t TMEM_COLS = DPT;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) init_ring_barriers(bars);
  if (tid < kThreads) {
    for (int i = tid; i < DPT; i += kThreads) { q_s[i] = 0.f; t1_s[i] = 0.f; }
    for (int i = tid; i < MP / KT; i += kThreads) {
      SlabDesc d;
      d.img = ZtTU + tc_slab_ztt(DPT, i);
      d.rows = DPT; d.tmem_off = 0; d.first = i == 0;
      tab[i] = d;
    }
  }
  if (tid < 4) sc_s[tid] = 0.f;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  const int tiles_mine = a.ntiles > (int)blockIdx.x ? (a.ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  if (warp == kIssuerWarp) {
    issuer_loop<DPT, true>(stage_base, bars, tmem_slot, tab, MP / KT, tiles_mine, blockIdx.x == 0 ? a.trace : nullptr, nullptr,
                           !(a.exp_mode & 8));
  } else {
  Pipe<DPT> pipe;
  pipe.init(stage_base, bars);
  long long dseg[6] = {0, 0, 0, 0, 0, 0};
  long long dlast = clock64();
#define DSEG(i) do { if (a.dbg) { const long long tnow = clock64(); dseg[i] += tnow - dlast; dlast = tnow; } } while (0)

  // epilogue mapping: the 8 warps cover 4 lane quadrants x 2 column halves of the [128, DPT] tile; with DPT = 32
  // only the first 4 warps have columns.
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;
  const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
  constexpr int CH_PER_HALF = DPT >= 64 ? DPT / 64 : 1;
  const bool has_cols = DPT >= 64 || half == 0;
  const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(a.dx) & 15) == 0);

  OpRegs<TNP> r0, r1, r2;
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const long long n0 = (long long)tile * TNP;
    auto load_w_at = [&](OpRegs<TNP>& regs, long long nbase, int sl) {
      load_kmajor<TNP>(regs, TNP, [&](int r, int c) {
        const long long gn = nbase + r;
        return gn < N ? ldg4(Wg + (size_t)gn * MP + sl * KT + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      });
    };
    auto load_w = [&](OpRegs<TNP>& regs, int sl) { load_w_at(regs, n0, sl); };
    // per-point scalars and the x rows of the first epilogue chunk: requested now, consumed after the GEMM
    const long long gn = n0 + row;
    const bool live = gn < N;
    const float r = live ? rrow[gn] : 0.f;
    const float gm = live ? gsc[gn] : 0.f;
    const float gv = live ? gsc[N + gn] : 0.f;
    const long long w0 = n0 + quad * 32;
    const int nvalid = N - w0 >= 32 ? 32 : (N > w0 ? (int)(N - w0) : 0);
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;
    auto load_x_chunk = [&](float4 (&xq)[8], int col) {
      const int d = col + c4;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int rr = it * 4 + rsub;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (rr < nvalid && d < D) {
          const float* xr = a.x + (size_t)(w0 + rr) * D + d;
          if (vec) xv = ldg4(xr);
          else {
            xv.x = xr[0];
            if (d + 1 < D) xv.y = xr[1];
            if (d + 2 < D) xv.z = xr[2];
            if (d + 3 < D) xv.w = xr[3];
          }
        }
        xq[it] = xv;
      }
    };
    float4 xq[8];
    if (has_cols) load_x_chunk(xq, DPT >= 64 ? half * (DPT / 2) : 0);
    // W slabs are fetched three slabs ahead (rotating register sets) to keep enough bytes in flight per SM; the
    // first three of a tile are requested before the PREVIOUS tile's epilogue (see below)
    const int nsl = MP / KT;
    DSEG(0);
    if (tile == (int)blockIdx.x) {
      load_w(r0, 0);
      if (1 < nsl) load_w(r1, 1);
      if (2 < nsl) load_w(r2, 2);
    }
    for (int s = 0; s < nsl; s += 3) {
      float *a_hi, *a_lo;
      long long* ptr = (a.trace && blockIdx.x == 0 && (tid == 0 || tid == 255) && pipe.slab < 48)
                           ? a.trace + 512 + (tid == 0 ? 0 : 256) + pipe.slab * 4 : nullptr;
      if (ptr) ptr[0] = clock64();
      pipe.acquire(a_hi, a_lo);
      if (ptr) ptr[1] = clock64();
      DSEG(1);
      store_kmajor<TNP>(a_hi, a_lo, r0, TNP);
      if (ptr) ptr[2] = clock64();
      DSEG(2);
      if (s + 3 < nsl) load_w(r0, s + 3);
      pipe.commit();
      if (ptr) ptr[3] = clock64();
      DSEG(3);
      if (s + 1 >= nsl) break;
      pipe.acquire(a_hi, a_lo);
      DSEG(1);
      store_kmajor<TNP>(a_hi, a_lo, r1, TNP);
      DSEG(2);
      if (s + 4 < nsl) load_w(r1, s + 4);
      pipe.commit();
      DSEG(3);
      if (s + 2 >= nsl) break;
      pipe.acquire(a_hi, a_lo);
      DSEG(1);
      store_kmajor<TNP>(a_hi, a_lo, r2, TNP);
      DSEG(2);
      if (s + 5 < nsl) load_w(r2, s + 5);
      pipe.commit();
      DSEG(3);
    }
    if (tile + (int)gridDim.x < a.ntiles) {         // next tile's first W slabs fly during this tile's epilogue
      const long long n1 = n0 + (long long)gridDim.x * TNP;
      load_w_at(r0, n1, 0);
      if (1 < nsl) load_w_at(r1, n1, 1);
      if (2 < nsl) load_w_at(r2, n1, 2);
    }
    pipe.drain();
    DSEG(4);

    if (has_cols) {
      // the warp owns the points n0 + quad * 32 + [0, 32) and 32 dimensions per chunk.  x comes in with coalesced
      // 16-byte loads (8 lanes per row segment), is centred / scaled once and parked in the warp's staging tile
      // red[warp][row][.] (pitch 36: conflict-free row AND column access); the per-dimension sums read it by column
      // (lane = dimension), the dx rows are written back through the same tile, transposed and coalesced.
      float (*stg)[kStagePitch] = red[warp];
      rg_s[warp][0][lane] = live ? r : 0.f;
      rg_s[warp][1][lane] = live ? gm : 0.f;
#pragma unroll 1
      for (int ch = 0; ch < CH_PER_HALF; ++ch) {
        const int col = (DPT >= 64 ? half * (DPT / 2) : 0) + ch * 32;
        __syncwarp();
        {
          const int d = col + c4;
          const float4 c4v = d < DP ? ldg4(center + d) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 iev = d < DP ? ldg4(inv_ell + d) : make_float4(0.f, 0.f, 0.f, 0.f);
          if (ch > 0) load_x_chunk(xq, col);           // DPT = 128 only; the first chunk was prefetched
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rsub;
            float4 xv = xq[it];
            if (rr < nvalid && d < D) {
              xv.x = (xv.x - c4v.x) * iev.x; xv.y = (xv.y - c4v.y) * iev.y;
              xv.z = (xv.z - c4v.z) * iev.z; xv.w = (xv.w - c4v.w) * iev.w;
            }
            *reinterpret_cast<float4*>(&stg[rr][c4]) = xv;
          }
        }
        __syncwarp();
        // per-dimension reductions over the 32 points of this warp (lane = dimension col + lane), fixed order:
        // q_d = sum r x~^2, t1_d = sum g_mu x~
        float sq = 0.f, st = 0.f;
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) {
          const float xv = stg[rr][lane];
          sq = fmaf(rg_s[warp][0][rr] * xv, xv, sq);
          st = fmaf(rg_s[warp][1][rr], xv, st);
        }
        // dx row of this lane's point: (W Z~ - r x~) / ell + g_mu w
        float v[32];
        tc::tmem_ld32(tmem_d + lane_base + (uint32_t)col, v);
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int d = col + i;
          const float4 xs4 = *reinterpret_cast<const float4*>(&stg[lane][i]);
          const float4 ie = d < DP ? ldg4(inv_ell + d) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 w4 = d < DP ? ldg4(wl + d) : make_float4(0.f, 0.f, 0.f, 0.f);
          v[i + 0] = (v[i + 0] - r * xs4.x) * ie.x + gm * (w4.x * ie.x);
          v[i + 1] = (v[i + 1] - r * xs4.y) * ie.y + gm * (w4.y * ie.y);
          v[i + 2] = (v[i + 2] - r * xs4.z) * ie.z + gm * (w4.z * ie.z);
          v[i + 3] = (v[i + 3] - r * xs4.w) * ie.w + gm * (w4.w * ie.w);
        }
        __syncwarp();                                // every lane is done reading x~ by column
        if (a.dx) {
#pragma unroll
          for (int i = 0; i < 32; i += 4)
            *reinterpret_cast<float4*>(&stg[lane][i]) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
          __syncwarp();
          const int d = col + c4;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rsub;
            if (rr < nvalid && d < D) {
              const float4 o = *reinterpret_cast<const float4*>(&stg[rr][c4]);
              float* dst = a.dx + (size_t)(w0 + rr) * D + d;
              if (vec) *reinterpret_cast<float4*>(dst) = o;
              else {
                dst[0] = o.x;
                if (d + 1 < D) dst[1] = o.y;
                if (d + 2 < D) dst[2] = o.z;
                if (d + 3 < D) dst[3] = o.w;
              }
            }
          }
        }
        // this warp's partial for dimension col + lane; (quad, half, ch) -> combined below in fixed order
        if (ch == 0) { part_q[warp][lane] = sq; part_t[warp][lane] = st; }
        else { part_q2[warp][lane] = sq; part_t2[warp][lane] = st; }   // DPT = 128: second chunk of the half
      }
    }
    {
      const float sg = warp_sum(live && half == 0 ? gm : 0.f);
      const float sr = warp_sum(live && half == 0 ? r : 0.f);
      const float sv = warp_sum(live && half == 0 ? gv : 0.f);
      if (lane == 0) { part_sc[warp][VS_GMU] = sg; part_sc[warp][VS_RSUM] = sr; part_sc[warp][VS_GVAR] = sv; }
    }
    prod_sync();
    // combine the 4 lane quadrants in fixed order: dimension d = half * (DPT / 2) + ch * 32 + lane
    if (tid < DPT) {
      const int d = tid;
      const int hf = DPT >= 64 ? d / (DPT / 2) : 0;
      const int within = DPT >= 64 ? d % (DPT / 2) : d;
      const int ch = within / 32, ln = within % 32;
      float sq = 0.f, st = 0.f;
      for (int qd = 0; qd < 4; ++qd) {
        const int w = hf * 4 + qd;
        if (ch == 0) { sq += part_q[w][ln]; st += part_t[w][ln]; }
        else { sq += part_q2[w][ln]; st += part_t2[w][ln]; }
      }
      q_s[d] += sq;
      t1_s[d] += st;
    }
    if (tid < 3) {
      float s = 0.f;
      for (int w = 0; w < 4; ++w) s += part_sc[w][tid];
      sc_s[tid] += s;
    }
    tc::tc_fence_before();   // the TMEM reads of this tile are done before the next tile's MMAs are released
    prod_sync();
    DSEG(5);
  }
  if (a.dbg && blockIdx.x == 0 && tid == 0)
    for (int i = 0; i < 6; ++i) a.dbg[24 + i] = dseg[i];
#undef DSEG
  }   // producer warps
  __syncthreads();
  float* vp = vecpart + (size_t)blockIdx.x * L.vec_len;
  if (tid < kThreads) {
    for (int i = tid; i < MP; i += kThreads) vp[i] = 0.f;        // column sums come from the W^T X kernel
    for (int d = tid; d < DP; d += kThreads) {
      vp[MP + d] = d < DPT ? q_s[d] : 0.f;
      vp[MP + DP + d] = (d < D && d < DPT) ? ellv[d] * t1_s[d] + center[d] * sc_s[VS_GMU] : 0.f;
    }
    if (tid < VS_COUNT) vp[MP + 2 * DP + tid] = tid < 3 ? sc_s[tid] : 0.f;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_slot, TMEM_COLS);
}

template <int NB>
constexpr size_t tc_smem_bytes() { return (size_t)2 * Stage<NB>::FLOATS * 4; }

int tc_grid(const WsLayout& L) {
  const long long nt = (L.N + TNP - 1) / TNP;
  const int sms = num_sms();
  return (int)(nt < sms ? (nt < 1 ? 1 : nt) : sms);
}

template <class K>
void set_smem(K kernel, size_t bytes) {
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

}  // namespace

bool tc_point_supported(const WsLayout& L) {
  if (tile_override("GPBLUR_TC") < 0) return false;       // GPBLUR_TC=-1 forces the FP32 FFMA kernels
  return (L.MP == 128 || (L.MP >= 256 && L.MP % 256 == 0)) && L.N >= 1;
}

int tc_vector_partials(const WsLayout& L) { return tc_grid(L); }

int launch_tc_point_forward(const WsLayout& L, void* ws, const float* x, float* mean, float* var, float* sample,
                            uint64_t seed, uint64_t offset, uint32_t stream_id, cudaStream_t st) {
  TcPointArgs a{};
  a.L = L; a.ws = ws; a.x = x; a.mean = mean; a.var = var; a.sample = sample;
  a.seed = seed; a.offset = offset; a.offset_dev = current_offset_dev(); a.stream_id = stream_id;
  a.ntiles = (int)((L.N + TNP - 1) / TNP);
  a.dbg = tile_override("GPBLUR_TC_DEBUG") > 0 ? ws_ptr<long long>(ws, L.stamps) : nullptr;
  a.exp_mode = tile_override("GPBLUR_TC_EXP");
  {
    const char* tp = getenv("GPBLUR_FWD_TRACE_PTR");
    a.trace = tp ? reinterpret_cast<long long*>(strtoull(tp, nullptr, 0)) : nullptr;
  }
  const int grid = tc_grid(L);
  ProfScope ps(ST_POINT_FWD, st);
  if (L.MP == 128) {
    static bool cfg = false;
    if (!cfg) { set_smem(tc_point_fwd_kernel<128>, tc_smem_bytes<128>()); cfg = true; }
    tc_point_fwd_kernel<128><<<grid, kTwoGroupThreads, tc_smem_bytes<128>(), st>>>(a);
  } else {
    static bool cfg = false;
    if (!cfg) { set_smem(tc_point_fwd_kernel<256>, tc_smem_bytes<256>()); cfg = true; }
    tc_point_fwd_kernel<256><<<grid, kTwoGroupThreads, tc_smem_bytes<256>(), st>>>(a);
  }
  note_launch();
  return check_launch("tc_point_fwd");
}

int launch_tc_point_backward(const WsLayout& L, void* ws, const float* x, const float* g_mean, const float* g_var,
                             const float* g_sample, const float* var, uint64_t seed, uint64_t offset,
                             uint32_t stream_id, float* dx, cudaStream_t st) {
  TcPointArgs a{};
  a.L = L; a.ws = ws; a.x = x; a.g_mean = g_mean; a.g_var = g_var; a.g_sample = g_sample; a.var_in = var; a.dx = dx;
  a.seed = seed; a.offset = offset; a.offset_dev = current_offset_dev(); a.stream_id = stream_id;
  a.ntiles = (int)((L.N + TNP - 1) / TNP);
  a.dbg = tile_override("GPBLUR_TC_DEBUG") > 0 ? ws_ptr<long long>(ws, L.stamps) : nullptr;
  a.exp_mode = tile_override("GPBLUR_TC_EXP");
  {
    const char* tp = getenv("GPBLUR_TRACE_PTR");
    a.trace = tp ? reinterpret_cast<long long*>(strtoull(tp, nullptr, 0)) : nullptr;
  }
  const int grid = tc_grid(L);
  {
    ProfScope ps(ST_POINT_BWD, st);
    if (L.MP == 128) {
      static bool cfg = false;
      if (!cfg) { set_smem(tc_point_bwd_kernel<128>, tc_smem_bytes<128>()); cfg = true; }
      tc_point_bwd_kernel<128><<<grid, kBwdThreads, tc_smem_bytes<128>(), st>>>(a);
    } else {
      static bool cfg = false;
      if (!cfg) { set_smem(tc_point_bwd_kernel<256>, tc_smem_bytes<256>()); cfg = true; }
      tc_point_bwd_kernel<256><<<grid, kBwdThreads, tc_smem_bytes<256>(), st>>>(a);
    }
    note_launch();
    int rc = check_launch("tc_point_bwd");
    if (rc) return rc;
  }
  ProfScope ps(ST_OTHER, st);
  if (L.DP <= 32) {
    static bool cfg = false;
    if (!cfg) { set_smem(tc_dx_kernel<32>, tc_smem_bytes<32>()); cfg = true; }
    tc_dx_kernel<32><<<grid, kBlockThreads, tc_smem_bytes<32>(), st>>>(a);
  } else if (L.DP == 64) {
    static bool cfg = false;
    if (!cfg) { set_smem(tc_dx_kernel<64>, tc_smem_bytes<64>()); cfg = true; }
    tc_dx_kernel<64><<<grid, kBlockThreads, tc_smem_bytes<64>(), st>>>(a);
  } else {
    static bool cfg = false;
    if (!cfg) { set_smem(tc_dx_kernel<128>, tc_smem_bytes<128>()); cfg = true; }
    tc_dx_kernel<128><<<grid, kBlockThreads, tc_smem_bytes<128>(), st>>>(a);
  }
  note_launch();
  return check_launch("tc_dx");
}

}  // namespace gpblur
