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
// Operand tiles are produced by all 256 threads (coalesced 16-byte loads, TF32 hi/lo split, canonical K-major
// no-swizzle UMMA layout), one elected thread issues the MMAs, tcgen05.commit -> mbarrier tracks completion.
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
  uint32_t stream_id;
  int ntiles;
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

// Two-stage operand ring shared by the producer warps (0..7) and the issuer warp (8).  Both roles run the SAME slab
// schedule (same loops, same counters); producers never wait for the issuer except through the barriers below:
//   bars[0..1] mma_done : tcgen05.commit - the MMAs that read stage st have retired (stage reusable)
//   bars[2..3] b_full   : the TMA bulk copy of stage st's B planes has landed (complete_tx)
//   bars[4..5] a_ready  : all 256 producers have written (and proxy-fenced) stage st's A planes
template <int NB>
struct Pipe {
  float* base;
  uint64_t* bars;
  uint32_t uses[2];
  int slab;
  bool prefetched;
  bool issuer;     // role of the calling warp
  bool elect;      // the one issuing thread (lane 0 of the issuer warp)
  long long iseg[6];   // issuer cycle accounting (debug)
  __device__ __forceinline__ void init(float* b, uint64_t* br) {
    base = b; bars = br; uses[0] = uses[1] = 0; slab = 0; prefetched = false;
    issuer = (threadIdx.x >> 5) == kIssuerWarp;
    elect = threadIdx.x == kThreads;
    for (int i = 0; i < 6; ++i) iseg[i] = 0;
  }
  __device__ __forceinline__ void planes(int st, float*& a_hi, float*& a_lo, float*& b_hi, float*& b_lo) const {
    a_hi = base + st * Stage<NB>::FLOATS;
    a_lo = a_hi + Stage<NB>::A_PLANE;
    b_hi = a_lo + Stage<NB>::A_PLANE;
    b_lo = b_hi + Stage<NB>::B_PLANE;
  }
  // producers: wait until the MMAs that last read this stage have retired, return its planes
  __device__ __forceinline__ void acquire(float*& a_hi, float*& a_lo, float*& b_hi, float*& b_lo) {
    const int st = slab & 1;
    if (!issuer && uses[st] > 0) tc::mbar_wait(&bars[st], (uses[st] - 1) & 1);
    planes(st, a_hi, a_lo, b_hi, b_lo);
  }
  __device__ __forceinline__ void issue_bulk(int st, const float* image, int rows) {
    float *a_hi, *a_lo, *b_hi, *b_lo;
    planes(st, a_hi, a_lo, b_hi, b_lo);
    const uint32_t bytes = (uint32_t)rows * 128u;          // 8 k-chunks x rows x 16 B
    tc::mbar_expect_tx(&bars[2 + st], 2 * bytes);
    tc::bulk_g2s(b_hi, image, bytes, &bars[2 + st]);
    tc::bulk_g2s(b_lo, image + (size_t)rows * 32, bytes, &bars[2 + st]);
  }
  // issuer: B operand of the CURRENT slab (a no-op when the previous commit() already prefetched it)
  __device__ __forceinline__ void bulk_b(const float* image, int rows) {
    if (elect && !prefetched) {
      const int st = slab & 1;
      if (uses[st] > 0) tc::mbar_wait(&bars[st], (uses[st] - 1) & 1);
      issue_bulk(st, image, rows);
    }
  }
  // producers: publish this stage's A planes (no CTA barrier).  issuer: wait for A and B, issue the N-column MMAs
  // into tmem_d, commit, then start the TMA copy of the NEXT slab's B image into the other stage as soon as the
  // MMAs that still read it have retired.
  __device__ __forceinline__ void commit(uint32_t tmem_d, int ncols, bool first, int b_rows, const float* next_image,
                                         int next_rows) {
    const int st = slab & 1;
    const uint32_t phase = uses[st] & 1;
    if (!issuer) {
      tc::tc_fence_before();
      tc::fence_async_smem();
      mbar_arrive(&bars[4 + st]);
    } else if (elect) {
      long long t0 = clock64(), t1;
      tc::mbar_wait(&bars[4 + st], phase);
      t1 = clock64(); iseg[0] += t1 - t0; t0 = t1;
      tc::mbar_wait(&bars[2 + st], phase);
      t1 = clock64(); iseg[1] += t1 - t0; t0 = t1;
      tc::tc_fence_after();
      float *a_hi, *a_lo, *b_hi, *b_lo;
      planes(st, a_hi, a_lo, b_hi, b_lo);
      tc::issue_slab_3xtf32<KT, NB>(tmem_d, a_hi, a_lo, b_hi, b_lo, tc::make_idesc_tf32(TNP, ncols), first, b_rows);
      tc::umma_commit(&bars[st]);
      t1 = clock64(); iseg[2] += t1 - t0; t0 = t1;
      if (next_image) {
        const int nst = st ^ 1;
        if (uses[nst] > 0) tc::mbar_wait(&bars[nst], (uses[nst] - 1) & 1);
        t1 = clock64(); iseg[3] += t1 - t0; t0 = t1;
        issue_bulk(nst, next_image, next_rows);
        t1 = clock64(); iseg[4] += t1 - t0; t0 = t1;
      }
    }
    uses[st] += 1;
    slab += 1;
    prefetched = next_image != nullptr;
  }
  // producers: block until every MMA issued so far has completed
  __device__ __forceinline__ void drain() {
    if (issuer || slab == 0) return;
    const int st = (slab - 1) & 1;
    tc::mbar_wait(&bars[st], (uses[st] - 1) & 1);
    tc::tc_fence_after();
  }
  __device__ __forceinline__ static void init_barriers(uint64_t* br) {
    tc::mbar_init(&br[0], 1); tc::mbar_init(&br[1], 1);
    tc::mbar_init(&br[2], 1); tc::mbar_init(&br[3], 1);
    tc::mbar_init(&br[4], kThreads); tc::mbar_init(&br[5], kThreads);
    tc::fence_barrier_init();
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
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rr = lane & 7, cq = lane >> 3;
  const int ntiles = (nrows >> 3) * 2;
#pragma unroll
  for (int p = 0; p < OpRegs<PLANE_ROWS>::PASSES; ++p) {
    const int wt = warp + 8 * p;
    if (wt < ntiles) tc::store_split(hi, lo, tc::op_off<PLANE_ROWS>((wt >> 1) * 8 + rr, (wt & 1) * 4 + cq), regs.v[p]);
  }
}

__device__ __forceinline__ float4 ldg4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// x tile row (point) -> centred / scaled float4 of k-chunk `dchunk`
struct XLoader {
  const float* x; long long n0, N; int D; const float* center; const float* inv_ell; bool vec;
  __device__ __forceinline__ float4 operator()(int row, int dchunk) const {
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
      const float4 c = ldg4(center + d), ie = ldg4(inv_ell + d);
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

// phase A of both kernels: S[128, MP] = X~ Z~^T into TMEM columns [0, MP)
// registers of one 32-wide d-slab of the x tile (centred / scaled)
__device__ __forceinline__ void load_x_slab(OpRegs<TNP>& ra, const XLoader& xl, int ds, int DP) {
  load_kmajor<TNP>(ra, TNP, [&](int row, int c) {
    const int dchunk = ds * (KT / 4) + c;
    return dchunk * 4 < DP ? xl(row, dchunk) : make_float4(0.f, 0.f, 0.f, 0.f);
  });
}

// S[128, BW] = X~ Z~[block q]^T into TMEM columns [0, BW).  xr0 / xr1: the first two d-slabs of this tile's x when
// `preloaded` (loaded one tile ahead so that the HBM latency hides behind the previous tile's MMAs and epilogue).
// With `stats`, per-row |x~|^2 and x~ . (ell w) are folded from per-slab partials in fixed order (deterministic).
template <int BW>
__device__ __forceinline__ void phase_a(Pipe<BW>& pipe, uint32_t tmem_s, const TcPointArgs& a, const XLoader& xl, int q,
                                        const float* after_image, int after_rows, bool stats, float* part_n,
                                        float* part_w, float* xn_s, float* xw_s, bool preloaded,
                                        const OpRegs<TNP>& xr0, const OpRegs<TNP>& xr1) {
  const WsLayout& L = a.L;
  const float* ZtU = ws_cptr<float>(a.ws, L.ZtU);
  const float* wl = ws_cptr<float>(a.ws, L.wl);
  const int DP = L.DP, MP = L.MP;
  const int nds = DP >= KT ? DP / KT : 1;
  const bool prod = !pipe.issuer;
  for (int ds = 0; ds < nds; ++ds) {
    OpRegs<TNP> ra;
    if (prod) {
      if (preloaded && ds == 0) ra = xr0;
      else if (preloaded && ds == 1) ra = xr1;
      else load_x_slab(ra, xl, ds, DP);
    }
    if (prod && stats) {
      // row-statistic partials of the chunks this thread just loaded (same (row, chunk) mapping as load_kmajor)
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
    float *a_hi, *a_lo, *b_hi, *b_lo;
    pipe.acquire(a_hi, a_lo, b_hi, b_lo);
    pipe.bulk_b(ZtU + tc_zt_image(MP, nds, q, ds), BW);   // Z~ image: TMA bulk copy of the pre-split block
    if (prod) store_kmajor<TNP>(a_hi, a_lo, ra, TNP);
    const bool last = ds + 1 == nds;
    pipe.commit(tmem_s, BW, ds == 0, BW, last ? after_image : ZtU + tc_zt_image(MP, nds, q, ds + 1),
                last ? after_rows : BW);
    if (prod && stats) {
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

// cross-covariance values of one 32-column chunk from the S accumulators: k = os exp(-1/2 max(|x|^2 + |z|^2 - 2 s, 0))
__device__ __forceinline__ void kernel_values(float (&v)[32], float xn, const float* zn_s, int col0, int M, float os) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const int m = col0 + i;
    const float d2 = fmaxf(xn + zn_s[m] - 2.0f * v[i], 0.f);
    v[i] = (m < M) ? os * tc::fast_exp(-0.5f * d2) : 0.f;
  }
}

// =================================================================================================
// forward.  The inducing dimension is processed in column blocks of width BW (= min(MP, 256), the TMEM budget:
// S in columns [0, BW), the whitened product of the current output block in [BW, 2 BW)).  Output block p needs the
// cross-covariance blocks q <= p (Linv is lower triangular); for MP > 256 block q is recomputed for every p >= q.
// =================================================================================================
template <int BW>
__global__ void __launch_bounds__(kBlockThreads, 1) tc_point_fwd_kernel(TcPointArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bars[6];
  __shared__ uint32_t tmem_slot;
  __shared__ float zn_s[GPBLUR_MAX_M], m_s[GPBLUR_MAX_M], c_s[GPBLUR_MAX_M];
  __shared__ float xn_s[TNP], xw_s[TNP], mu_s[TNP], vv_s[TNP];
  __shared__ float part_n[(KT / 4) * TNP], part_w[(KT / 4) * TNP];   // per-slab row-statistic partials

  const WsLayout& L = a.L;
  const int M = L.M, MP = L.MP;
  const int NP = MP / BW, SPB = BW / KT;
  const long long N = L.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* hyp = ws_cptr<float>(a.ws, L.hyp);
  const float* LinvU = ws_cptr<float>(a.ws, L.LinvU);
  const float* ZtU = ws_cptr<float>(a.ws, L.ZtU);
  float* Ag = ws_ptr<float>(a.ws, L.A);
  const float os = hyp[H_OS], jit = hyp[H_JIT], cwb = hyp[H_CWB];
  const int nds = L.DP >= KT ? L.DP / KT : 1;

  constexpr uint32_t TMEM_COLS = 2 * BW;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) Pipe<1>::init_barriers(bars);
  if (tid < kThreads)
    for (int i = tid; i < MP; i += kThreads) {
      zn_s[i] = ws_cptr<float>(a.ws, L.zn)[i];
      m_s[i] = ws_cptr<float>(a.ws, L.mvec)[i];
      c_s[i] = ws_cptr<float>(a.ws, L.cvec)[i];
    }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_s = tmem_slot, tmem_a = tmem_slot + BW;
  Pipe<BW> pipe;
  pipe.init(stage_base, bars);
  const bool prod = !pipe.issuer;                   // warps 0..7 produce operands and run the epilogues

  const int quad = warp & 3, half = warp >> 2;      // TMEM lane quadrant / column half of this warp
  const int row = quad * 32 + lane;                 // the point this thread owns in the epilogues
  const uint32_t lane_base = (uint32_t)(quad * 32) << 16;

  long long seg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = clock64();
#define SEG(i) do { if (a.dbg) { const long long tnow = clock64(); seg[i] += tnow - tlast; tlast = tnow; } } while (0)
  OpRegs<TNP> xr0, xr1;
  if (prod) {
    const XLoader xl0 = make_xloader(a, (long long)blockIdx.x * TNP);
    load_x_slab(xr0, xl0, 0, L.DP);
    if (L.DP > KT) load_x_slab(xr1, xl0, 1, L.DP);
  }
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const long long n0 = (long long)tile * TNP;
    const long long gn = n0 + row;
    const XLoader xl = make_xloader(a, n0);
    const bool more_tiles = tile + (int)gridDim.x < a.ntiles;
    float mu = 0.f, vv = 0.f;
    SEG(7);
    for (int p = 0; p < NP; ++p) {
      for (int q = 0; q <= p; ++q) {
        int rows0;
        const float* first_linv = LinvU + tc_linv_image(MP, p, q * SPB, &rows0);
        const bool first_pass = p == 0 && q == 0;
        phase_a<BW>(pipe, tmem_s, a, xl, q, first_linv, rows0, first_pass, part_n, part_w, xn_s, xw_s, first_pass, xr0,
                    xr1);
        if (prod && first_pass && more_tiles) {     // next tile's x: in flight during the MMAs and epilogues
          const XLoader xln = make_xloader(a, n0 + (long long)gridDim.x * TNP);
          load_x_slab(xr0, xln, 0, L.DP);
          if (L.DP > KT) load_x_slab(xr1, xln, 1, L.DP);
        }
        SEG(0);                                     // phase A (x split, Z~ bulk, S MMAs issued)
        pipe.drain();                               // S of block q complete
        SEG(1);
        const float xn = prod ? xn_s[row] : 0.f;
        // ---- whitening: A[:, block p] += k[:, slab s] Linv[block p rows >= 32 s, slab s]^T ----
        for (int sl = 0; sl < SPB; ++sl) {
          const int s = q * SPB + sl;               // global k-slab
          // fused epilogue of S: the two column halves of a lane quadrant take 16 columns each
          float v[16];
          const int col0 = sl * KT + half * 16;
          if (prod) {
            tc::tmem_ld16(tmem_s + lane_base + (uint32_t)col0, v);
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int m = q * BW + col0 + i;
              const float d2 = fmaxf(xn + zn_s[m] - 2.0f * v[i], 0.f);
              v[i] = (m < M) ? os * tc::fast_exp(-0.5f * d2) : 0.f;
            }
          }
          SEG(2);                                   // TMEM load + exp
          int rows;
          const float* img = LinvU + tc_linv_image(MP, p, s, &rows);
          float *a_hi, *a_lo, *b_hi, *b_lo;
          pipe.acquire(a_hi, a_lo, b_hi, b_lo);
          SEG(3);                                   // stage acquire (MMA s-2 retired)
          pipe.bulk_b(img, rows);
          if (prod) {
#pragma unroll
            for (int c = 0; c < 4; ++c)             // k-chunks half * 4 + c of the slab
              tc::store_split(a_hi, a_lo, tc::op_off<TNP>(row, half * 4 + c),
                              make_float4(v[c * 4 + 0], v[c * 4 + 1], v[c * 4 + 2], v[c * 4 + 3]));
          }
          // what to prefetch next: the next Linv slab of this block, else the Z~ image of the next (p, q) / tile
          const float* nxt = nullptr;
          int nxt_rows = BW;
          if (sl + 1 < SPB) nxt = LinvU + tc_linv_image(MP, p, s + 1, &nxt_rows);
          else if (q < p) nxt = ZtU + tc_zt_image(MP, nds, q + 1, 0);
          else if (p + 1 < NP || more_tiles) nxt = ZtU + tc_zt_image(MP, nds, 0, 0);
          SEG(4);                                   // split + store of the k slab
          pipe.commit(tmem_a + (uint32_t)(BW - rows), rows, q == 0 && sl == 0, rows, nxt, nxt_rows);
          SEG(5);                                   // fence + barrier (+ thread 0: wait B, issue MMAs, prefetch)
        }
      }
      pipe.drain();
      SEG(1);
      // ---- epilogue of output block p: mean / variance partials of the own point; save A ----
      if (prod) {
#pragma unroll 1
      for (int ch = 0; ch < BW / 64; ++ch) {
        const int col = half * (BW / 2) + ch * 32;
        float v[32];
        tc::tmem_ld32(tmem_a + lane_base + (uint32_t)col, v);
        const int gcol = p * BW + col;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          mu = fmaf(v[i], m_s[gcol + i], mu);
          vv = fmaf(c_s[gcol + i] * v[i], v[i], vv);
        }
        if (L.training && gn < N) {
          float* dst = Ag + (size_t)gn * MP + gcol;
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        }
      }
      }   // the next MMAs into these columns are issued only after every producer's next commit(), which fences
      SEG(6);                                       // block epilogue
    }
    if (prod) {
    if (half == 1) { mu_s[row] = mu; vv_s[row] = vv; }
    prod_sync();
    if (half == 0 && gn < N) {
      const float mean = mu + mu_s[row] + xw_s[row] + cwb;
      const float var = fmaxf(os + jit + vv + vv_s[row], kMinVariance);
      a.mean[gn] = mean;
      a.var[gn] = var;
      if (a.sample) a.sample[gn] = fmaf(sqrtf(var), philox_normal(a.seed, a.offset + (uint64_t)gn, a.stream_id), mean);
    }
    prod_sync();
    }
  }
  if (a.dbg && blockIdx.x == 0 && (tid == 0 || tid == 32))
    for (int i = 0; i < 8; ++i) a.dbg[(tid == 0 ? 0 : 8) + i] = seg[i];
  if (a.dbg && blockIdx.x == 0 && pipe.elect)
    for (int i = 0; i < 6; ++i) a.dbg[16 + i] = pipe.iseg[i];
#undef SEG
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem_slot, TMEM_COLS);
}

// =================================================================================================
// backward: W = kbar o k and its row sums, one column block p of width BW at a time
// =================================================================================================
template <int BW>
__global__ void __launch_bounds__(kBlockThreads, 1) tc_point_bwd_kernel(TcPointArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bars[6];
  __shared__ uint32_t tmem_slot;
  __shared__ float zn_s[GPBLUR_MAX_M], beta_s[GPBLUR_MAX_M];
  __shared__ float xn_s[TNP], xw_s[TNP], r_s[TNP];
  __shared__ float part_n[(KT / 4) * TNP], part_w[(KT / 4) * TNP];

  const WsLayout& L = a.L;
  const int M = L.M, MP = L.MP;
  const int NP = MP / BW, SPB = BW / KT, NSL = MP / KT;
  const long long N = L.N;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float* hyp = ws_cptr<float>(a.ws, L.hyp);
  const float* LCTU = ws_cptr<float>(a.ws, L.LCTU);
  const float* ZtU = ws_cptr<float>(a.ws, L.ZtU);
  const float* Ag = ws_cptr<float>(a.ws, L.A);
  float* Wg = ws_ptr<float>(a.ws, L.W);
  float* gsc = ws_ptr<float>(a.ws, L.gsc);
  float* rrow = ws_ptr<float>(a.ws, L.rrow);
  const float os = hyp[H_OS];
  const int nds = L.DP >= KT ? L.DP / KT : 1;

  constexpr uint32_t TMEM_COLS = 2 * BW;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, TMEM_COLS);
  if (tid == 0) Pipe<1>::init_barriers(bars);
  if (tid < kThreads)
    for (int i = tid; i < MP; i += kThreads) {
      zn_s[i] = ws_cptr<float>(a.ws, L.zn)[i];
      beta_s[i] = ws_cptr<float>(a.ws, L.beta)[i];
    }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_s = tmem_slot, tmem_t = tmem_slot + BW;
  Pipe<BW> pipe;
  pipe.init(stage_base, bars);
  const bool prod = !pipe.issuer;

  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;
  const uint32_t lane_base = (uint32_t)(quad * 32) << 16;

  OpRegs<TNP> xr0, xr1;
  if (prod) {
    const XLoader xl0 = make_xloader(a, (long long)blockIdx.x * TNP);
    load_x_slab(xr0, xl0, 0, L.DP);
    if (L.DP > KT) load_x_slab(xr1, xl0, 1, L.DP);
  }
  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const long long n0 = (long long)tile * TNP;
    const long long gn = n0 + row;
    const XLoader xl = make_xloader(a, n0);
    const bool more_tiles = tile + (int)gridDim.x < a.ntiles;
    // ---- fold the upstream gradients of this thread's point ----
    float gm = 0.f, gv = 0.f;
    if (prod && gn < N) {
      if (a.g_mean) gm = a.g_mean[gn];
      if (a.g_var) gv = a.g_var[gn];
      const float v = a.var_in[gn];
      if (a.g_sample) {
        const float gs = a.g_sample[gn];
        const float eps = philox_normal(a.seed, a.offset + (uint64_t)gn, a.stream_id);
        gm += gs;
        gv = fmaf(gs * eps, 0.5f * rsqrtf(v), gv);
      }
      if (v <= kMinVariance) gv = 0.f;
      if (half == 0) { gsc[gn] = gm; gsc[N + gn] = gv; }
    }
    auto load_a = [&](OpRegs<TNP>& regs, int sl) {
      load_kmajor<TNP>(regs, TNP, [&](int r, int c) {
        long long g2 = n0 + r;
        if (g2 >= N) g2 = N - 1;                       // clamped rows carry g = 0
        return ldg4(Ag + (size_t)g2 * MP + sl * KT + c * 4);
      });
    };
    float rsum = 0.f;
    for (int p = 0; p < NP; ++p) {
      int rows_top;
      const float* top_img = LCTU + tc_lct_image(MP, p, NSL - 1, &rows_top);
      phase_a<BW>(pipe, tmem_s, a, xl, p, top_img, rows_top, p == 0, part_n, part_w, xn_s, xw_s, p == 0, xr0, xr1);
      if (prod && p == 0 && more_tiles) {
        const XLoader xln = make_xloader(a, n0 + (long long)gridDim.x * TNP);
        load_x_slab(xr0, xln, 0, L.DP);
        if (L.DP > KT) load_x_slab(xr1, xln, 1, L.DP);
      }
      // ---- T[:, block p] += a[:, slab s] (diag(c) Linv)[slab s, block p], slabs in DEcreasing order ----
      OpRegs<TNP> ra;
      if (prod) load_a(ra, NSL - 1);
      for (int s = NSL - 1; s >= p * SPB; --s) {
        int rows;
        const float* img = LCTU + tc_lct_image(MP, p, s, &rows);
        float *a_hi, *a_lo, *b_hi, *b_lo;
        pipe.acquire(a_hi, a_lo, b_hi, b_lo);
        pipe.bulk_b(img, rows);
        if (prod) store_kmajor<TNP>(a_hi, a_lo, ra, TNP);
        const float* nxt = nullptr;
        int nxt_rows = BW;
        if (s > p * SPB) {
          if (prod) load_a(ra, s - 1);                // next slab's saved-A tile flies during the MMAs
          nxt = LCTU + tc_lct_image(MP, p, s - 1, &nxt_rows);
        } else if (p + 1 < NP) {
          nxt = ZtU + tc_zt_image(MP, nds, p + 1, 0);
        } else if (more_tiles) {
          nxt = ZtU + tc_zt_image(MP, nds, 0, 0);
        }
        pipe.commit(tmem_t, rows, s == NSL - 1, rows, nxt, nxt_rows);
      }
      pipe.drain();
      if (prod) {
      const float xn = xn_s[row];
#pragma unroll 1
      for (int ch = 0; ch < BW / 64; ++ch) {
        const int col = half * (BW / 2) + ch * 32;
        const int gcol = p * BW + col;
        float k[32], t[32];
        tc::tmem_ld32(tmem_s + lane_base + (uint32_t)col, k);
        tc::tmem_ld32(tmem_t + lane_base + (uint32_t)col, t);
        kernel_values(k, xn, zn_s, gcol, M, os);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const float kb = fmaf(2.0f * gv, t[i], gm * beta_s[gcol + i]);
          t[i] = kb * k[i];
          rsum += t[i];
        }
        if (gn < N) {
          float* dst = Wg + (size_t)gn * MP + gcol;
#pragma unroll
          for (int i = 0; i < 32; i += 4) *reinterpret_cast<float4*>(dst + i) = make_float4(t[i], t[i + 1], t[i + 2], t[i + 3]);
        }
      }
      }
    }
    if (prod) {
      if (half == 1) r_s[row] = rsum;
      prod_sync();
      if (half == 0 && gn < N) rrow[gn] = rsum + r_s[row];
      prod_sync();
    }
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
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  float* stage_base = reinterpret_cast<float*>(smem_raw);
  __shared__ __align__(8) uint64_t bars[6];
  __shared__ uint32_t tmem_slot;
  __shared__ float red[8][32][33];
  __shared__ float part_q[8][32], part_t[8][32], part_sc[8][4];
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
  if (tid == 0) Pipe<1>::init_barriers(bars);
  if (tid < kThreads)
    for (int i = tid; i < DPT; i += kThreads) { q_s[i] = 0.f; t1_s[i] = 0.f; }
  if (tid < 4) sc_s[tid] = 0.f;
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem_d = tmem_slot;
  Pipe<DPT> pipe;
  pipe.init(stage_base, bars);
  const bool prod = !pipe.issuer;

  // epilogue mapping: the 8 warps cover 4 lane quadrants x 2 column halves of the [128, DPT] tile; with DPT = 32
  // only the first 4 warps have columns.
  const int quad = warp & 3, half = warp >> 2;
  const int row = quad * 32 + lane;
  const uint32_t lane_base = (uint32_t)(quad * 32) << 16;
  constexpr int CH_PER_HALF = DPT >= 64 ? DPT / 64 : 1;
  const bool has_cols = DPT >= 64 || half == 0;
  const bool vec = (D % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15) == 0) &&
                   ((reinterpret_cast<uintptr_t>(a.dx) & 15) == 0);

  for (int tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
    const long long n0 = (long long)tile * TNP;
    auto load_w = [&](OpRegs<TNP>& regs, int sl) {
      load_kmajor<TNP>(regs, TNP, [&](int r, int c) {
        const long long gn = n0 + r;
        return gn < N ? ldg4(Wg + (size_t)gn * MP + sl * KT + c * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
      });
    };
    OpRegs<TNP> ra;
    if (prod) load_w(ra, 0);
    for (int s = 0; s < MP / KT; ++s) {
      float *a_hi, *a_lo, *b_hi, *b_lo;
      pipe.acquire(a_hi, a_lo, b_hi, b_lo);
      pipe.bulk_b(ZtTU + tc_slab_ztt(DPT, s), DPT);
      if (prod) store_kmajor<TNP>(a_hi, a_lo, ra, TNP);
      if (prod && s + 1 < MP / KT) load_w(ra, s + 1);
      const bool last = s + 1 == MP / KT;
      const bool more_tiles = tile + (int)gridDim.x < a.ntiles;
      const float* nxt = last ? (more_tiles ? ZtTU : nullptr) : ZtTU + tc_slab_ztt(DPT, s + 1);
      pipe.commit(tmem_d, DPT, s == 0, DPT, nxt, DPT);
    }
    pipe.drain();
    if (!prod) continue;                            // the issuer warp only runs the slab schedule

    const long long gn = n0 + row;
    const bool live = gn < N;
    const float r = live ? rrow[gn] : 0.f;
    const float gm = live ? gsc[gn] : 0.f;
    const float gv = live ? gsc[N + gn] : 0.f;
    if (has_cols) {
#pragma unroll 1
      for (int ch = 0; ch < CH_PER_HALF; ++ch) {
        const int col = (DPT >= 64 ? half * (DPT / 2) : 0) + ch * 32;
        float v[32];
        tc::tmem_ld32(tmem_d + lane_base + (uint32_t)col, v);
        float xs[32];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const int d = col + i;
          float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
          if (live && d < D) {
            const float* xr = a.x + (size_t)gn * D;
            if (vec) xv = ldg4(xr + d);
            else {
              xv.x = xr[d];
              if (d + 1 < D) xv.y = xr[d + 1];
              if (d + 2 < D) xv.z = xr[d + 2];
              if (d + 3 < D) xv.w = xr[d + 3];
            }
          }
          const float4 c4 = d < DP ? ldg4(center + d) : make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 ie = d < DP ? ldg4(inv_ell + d) : make_float4(0.f, 0.f, 0.f, 0.f);
          xs[i + 0] = (xv.x - c4.x) * ie.x; xs[i + 1] = (xv.y - c4.y) * ie.y;
          xs[i + 2] = (xv.z - c4.z) * ie.z; xs[i + 3] = (xv.w - c4.w) * ie.w;
        }
        // dx
        if (a.dx && live) {
          float* dst = a.dx + (size_t)gn * D + col;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            const int d = col + i;
            if (d < D) {
              const float4 ie = ldg4(inv_ell + d), w4 = ldg4(wl + d);
              float4 o;
              o.x = (v[i + 0] - r * xs[i + 0]) * ie.x + gm * (w4.x * ie.x);
              o.y = (v[i + 1] - r * xs[i + 1]) * ie.y + gm * (w4.y * ie.y);
              o.z = (v[i + 2] - r * xs[i + 2]) * ie.z + gm * (w4.z * ie.z);
              o.w = (v[i + 3] - r * xs[i + 3]) * ie.w + gm * (w4.w * ie.w);
              if (vec) *reinterpret_cast<float4*>(dst + i) = o;
              else {
                dst[i] = o.x;
                if (d + 1 < D) dst[i + 1] = o.y;
                if (d + 2 < D) dst[i + 2] = o.z;
                if (d + 3 < D) dst[i + 3] = o.w;
              }
            }
          }
        }
        // per-dimension reductions over the 32 points of this warp: q_d = sum r x~^2, t1_d = sum g_mu x~
#pragma unroll
        for (int i = 0; i < 32; ++i) red[warp][lane][i] = live ? r * xs[i] * xs[i] : 0.f;
        __syncwarp();
        float sq = 0.f;
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) sq += red[warp][rr][lane];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 32; ++i) red[warp][lane][i] = live ? gm * xs[i] : 0.f;
        __syncwarp();
        float st = 0.f;
#pragma unroll 8
        for (int rr = 0; rr < 32; ++rr) st += red[warp][rr][lane];
        __syncwarp();
        // this warp's partial for dimension col + lane; (quad, half, ch) -> combined below in fixed order
        if (CH_PER_HALF == 1) { part_q[warp][lane] = sq; part_t[warp][lane] = st; }
        else {
          // DPT = 128: two chunks per half; accumulate the second chunk into a second slot of the same warp
          if (ch == 0) { part_q[warp][lane] = sq; part_t[warp][lane] = st; }
          else { red[warp][0][lane] = sq; red[warp][1][lane] = st; }
        }
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
        else { sq += red[w][0][ln]; st += red[w][1][ln]; }
      }
      q_s[d] += sq;
      t1_s[d] += st;
    }
    if (tid < 3) {
      float s = 0.f;
      for (int w = 0; w < 4; ++w) s += part_sc[w][tid];
      sc_s[tid] += s;
    }
    prod_sync();
  }
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
constexpr size_t tc_smem_bytes() { return (size_t)2 * Stage<NB>::FLOATS * 4 + 1024; }

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
  a.seed = seed; a.offset = offset; a.stream_id = stream_id;
  a.ntiles = (int)((L.N + TNP - 1) / TNP);
  a.dbg = tile_override("GPBLUR_TC_DEBUG") > 0 ? ws_ptr<long long>(ws, L.stamps) : nullptr;
  const int grid = tc_grid(L);
  ProfScope ps(ST_POINT_FWD, st);
  if (L.MP == 128) {
    static bool cfg = false;
    if (!cfg) { set_smem(tc_point_fwd_kernel<128>, tc_smem_bytes<128>()); cfg = true; }
    tc_point_fwd_kernel<128><<<grid, kBlockThreads, tc_smem_bytes<128>(), st>>>(a);
  } else {
    static bool cfg = false;
    if (!cfg) { set_smem(tc_point_fwd_kernel<256>, tc_smem_bytes<256>()); cfg = true; }
    tc_point_fwd_kernel<256><<<grid, kBlockThreads, tc_smem_bytes<256>(), st>>>(a);
  }
  note_launch();
  return check_launch("tc_point_fwd");
}

int launch_tc_point_backward(const WsLayout& L, void* ws, const float* x, const float* g_mean, const float* g_var,
                             const float* g_sample, const float* var, uint64_t seed, uint64_t offset,
                             uint32_t stream_id, float* dx, cudaStream_t st) {
  TcPointArgs a{};
  a.L = L; a.ws = ws; a.x = x; a.g_mean = g_mean; a.g_var = g_var; a.g_sample = g_sample; a.var_in = var; a.dx = dx;
  a.seed = seed; a.offset = offset; a.stream_id = stream_id;
  a.ntiles = (int)((L.N + TNP - 1) / TNP);
  const int grid = tc_grid(L);
  {
    ProfScope ps(ST_POINT_BWD, st);
    if (L.MP == 128) {
      static bool cfg = false;
      if (!cfg) { set_smem(tc_point_bwd_kernel<128>, tc_smem_bytes<128>()); cfg = true; }
      tc_point_bwd_kernel<128><<<grid, kBlockThreads, tc_smem_bytes<128>(), st>>>(a);
    } else {
      static bool cfg = false;
      if (!cfg) { set_smem(tc_point_bwd_kernel<256>, tc_smem_bytes<256>()); cfg = true; }
      tc_point_bwd_kernel<256><<<grid, kBlockThreads, tc_smem_bytes<256>(), st>>>(a);
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
