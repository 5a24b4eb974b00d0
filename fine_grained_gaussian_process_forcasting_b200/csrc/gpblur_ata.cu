// Core of the ATA attention head as ONE pass forward and ONE backward.
//
// Reference (/root/reference/forecasting_models/ATA.py:53-65, called from modules/multi_head_attention.py:49-51):
//   Q_proj = Q_p.reshape(b, h, l, -1);   Q, _ = torch.topk(Q_proj, dim=-1, k=1)           -> q [b, h, l, 1]
//   K_proj = K_p.reshape(b, h, l_k, -1); K, _ = torch.topk(K_proj, dim=-1, k=1)           -> k [b, h, l_k, 1]
//   scores = einsum('bhqd,bhkd->bhqk', Q, K) / sqrt(d_k);  attn = softmax(scores, -1)
//   context = einsum('bhqk,bhkd->bhqd', attn, V)
// After the top-1 pooling the score matrix is RANK ONE (s_ij = q_i k_j / sqrt(d_k)), yet the reference materialises
// scores and attn [b, h, l, l_k] in HBM (302 MB each at b = 256, h = 8, l = l_k = 192) several times over (einsum,
// division, softmax, einsum, and again in the backward).  Here one CTA owns a (batch, head) pair: the pooled keys and
// the value rows sit in shared memory, a thread owns a query row and streams over the keys with the softmax
// normaliser and the context accumulators in registers; nothing of size l x l_k ever exists.  HBM traffic = the
// inputs once + the context: the kernel is bound by exp / FMA issue on the CUDA cores, not by memory (it is
// not GEMM-shaped: d = 1 on the score side, d_v = 4 ... 16 on the value side).
//
// Backward (what autograd does through softmax / einsum / topk in the reference): with p_ij = exp(a_i k_j - lse_i),
// a_i = q_i / sqrt(d_k), D_i = g_i . ctx_i:   gs_ij = p_ij (g_i . V_j - D_i)
//   g_q_i = sum_j gs_ij k_j / sqrt(d_k)      g_k_j = sum_i gs_ij a_i      g_V_j = sum_i p_ij g_i
// and the pooled gradients go to the arg-max element of their group (topk's backward), zeros elsewhere.  Phase 1 is
// thread-per-query, phase 2 thread-per-key over the query-side vectors in shared memory: every output element is
// produced by one thread in a fixed order (bit-deterministic).
#include "gpblur_common.cuh"

namespace gpblur {

namespace {

constexpr int kAtaThreads = 256;

struct AtaArgs {
  const float* qp;      // [B, H, Lq, G]
  const float* kp;      // [B, H, Lk, G]
  const float* v;       // element (b, h, j, e) at v + b v_sb + h v_sh + j v_sl + e
  long long v_sb, v_sh, v_sl;
  int B, H, Lq, Lk, G, DV;
  // nf > 0: qp / kp (and their gradients) are [B, nf * C, L] buffers - the output of ONE convolution that holds the
  // reference's nf filter stacks as channel groups - instead of the reference's torch.cat(dim=0) of nf [B, C, L]
  // tensors.  The kernel then reads group f0 of the cat order at its place in that buffer: block (filter i, batch j) of
  // blk = C * L floats sits at (j * nf + i) * blk.  Requires blk % G == 0 (a group never straddles two blocks).
  int nf;
  long long blk_q, blk_k;
  float scale;          // 1 / sqrt(d_k)
  float* ctx;           // [B, Lq, H, DV]
  float* q_pool;        // [B, H, Lq]
  float* k_pool;        // [B, H, Lk]
  int* q_arg;
  int* k_arg;
  float* lse;           // [B, H, Lq]
  // backward
  const float* g_ctx;   // [B, Lq, H, DV]
  float* g_qp;          // [B, H, Lq, G]
  float* g_kp;          // [B, H, Lk, G]
  float* g_v;           // [B, Lk, H, DV]
};

// float offset of the group that starts at cat-order offset f0 (see AtaArgs::nf)
__device__ __forceinline__ size_t group_offset(size_t f0, int nf, long long blk, int B) {
  if (nf == 0) return f0;
  const size_t block = f0 / (size_t)blk, rem = f0 - block * (size_t)blk;
  const size_t i = block / (size_t)B, j = block - i * (size_t)B;
  return (j * (size_t)nf + i) * (size_t)blk + rem;
}

// top-1 of a group of G floats (first maximum wins, as a sequential scan does)
__device__ __forceinline__ float pool_group(const float* __restrict__ row, int G, int& arg) {
  float best = row[0];
  int a = 0;
  if ((G & 3) == 0) {
    for (int g = 0; g < G; g += 4) {
      const float4 v = *reinterpret_cast<const float4*>(row + g);
      if (v.x > best) { best = v.x; a = g; }
      if (v.y > best) { best = v.y; a = g + 1; }
      if (v.z > best) { best = v.z; a = g + 2; }
      if (v.w > best) { best = v.w; a = g + 3; }
    }
  } else {
    for (int g = 1; g < G; ++g) {
      const float v = row[g];
      if (v > best) { best = v; a = g; }
    }
  }
  arg = a;
  return best;
}

__device__ __forceinline__ void scatter_group(float* __restrict__ row, int G, int arg, float val) {
  if ((G & 3) == 0) {
    for (int g = 0; g < G; g += 4) {
      float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
      if (arg == g) o.x = val;
      if (arg == g + 1) o.y = val;
      if (arg == g + 2) o.z = val;
      if (arg == g + 3) o.w = val;
      *reinterpret_cast<float4*>(row + g) = o;
    }
  } else {
    for (int g = 0; g < G; ++g) row[g] = (g == arg) ? val : 0.f;
  }
}

template <int DVP>
__global__ void __launch_bounds__(kAtaThreads) ata_fwd_kernel(AtaArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int Lk = a.Lk, Lq = a.Lq, G = a.G, DV = a.DV;
  float* ks = sm;                            // [Lk]
  float* vs = sm + round_up(Lk, 4);          // [Lk][DVP]
  const int bh = blockIdx.x, b = bh / a.H, h = bh - b * a.H;
  for (int j = threadIdx.x; j < Lk; j += kAtaThreads) {
    int arg;
    const float kv = pool_group(a.kp + group_offset(((size_t)bh * Lk + j) * G, a.nf, a.blk_k, a.B), G, arg);
    ks[j] = kv;
    a.k_pool[(size_t)bh * Lk + j] = kv;
    a.k_arg[(size_t)bh * Lk + j] = arg;
    const float* vr = a.v + b * a.v_sb + h * a.v_sh + j * a.v_sl;
#pragma unroll
    for (int e = 0; e < DVP; ++e) vs[j * DVP + e] = e < DV ? vr[e] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < Lq; i += kAtaThreads) {
    int arg;
    const float q = pool_group(a.qp + group_offset(((size_t)bh * Lq + i) * G, a.nf, a.blk_q, a.B), G, arg);
    a.q_pool[(size_t)bh * Lq + i] = q;
    a.q_arg[(size_t)bh * Lq + i] = arg;
    const float aq = q * a.scale;
    float m = -INFINITY;
    for (int j = 0; j < Lk; ++j) m = fmaxf(m, aq * ks[j]);
    float Z = 0.f, acc[DVP];
#pragma unroll
    for (int e = 0; e < DVP; ++e) acc[e] = 0.f;
    for (int j = 0; j < Lk; ++j) {
      const float p = __expf(aq * ks[j] - m);
      Z += p;
#pragma unroll
      for (int e = 0; e < DVP; ++e) acc[e] = fmaf(p, vs[j * DVP + e], acc[e]);
    }
    const float rz = 1.0f / Z;
    float* o = a.ctx + (((size_t)b * Lq + i) * a.H + h) * DV;
#pragma unroll
    for (int e = 0; e < DVP; ++e)
      if (e < DV) o[e] = acc[e] * rz;
    a.lse[(size_t)bh * Lq + i] = m + __logf(Z);
  }
}

template <int DVP>
__global__ void __launch_bounds__(kAtaThreads) ata_bwd_kernel(AtaArgs a) {
  extern __shared__ __align__(16) float sm[];
  const int Lk = a.Lk, Lq = a.Lq, G = a.G, DV = a.DV;
  float* ks = sm;                                  // [Lk]
  float* vs = ks + round_up(Lk, 4);                // [Lk][DVP]
  float* aqs = vs + (size_t)Lk * DVP;              // [Lq]  a_i
  float* lses = aqs + round_up(Lq, 4);             // [Lq]
  float* Ds = lses + round_up(Lq, 4);              // [Lq]  g_i . ctx_i
  float* gs = Ds + round_up(Lq, 4);                // [Lq][DVP]
  const int bh = blockIdx.x, b = bh / a.H, h = bh - b * a.H;
  for (int j = threadIdx.x; j < Lk; j += kAtaThreads) {
    ks[j] = a.k_pool[(size_t)bh * Lk + j];
    const float* vr = a.v + b * a.v_sb + h * a.v_sh + j * a.v_sl;
#pragma unroll
    for (int e = 0; e < DVP; ++e) vs[j * DVP + e] = e < DV ? vr[e] : 0.f;
  }
  for (int i = threadIdx.x; i < Lq; i += kAtaThreads) {
    const size_t o = (((size_t)b * Lq + i) * a.H + h) * DV;
    float d = 0.f;
#pragma unroll
    for (int e = 0; e < DVP; ++e) {
      const float g = e < DV ? a.g_ctx[o + e] : 0.f;
      gs[i * DVP + e] = g;
      if (e < DV) d = fmaf(g, a.ctx[o + e], d);
    }
    Ds[i] = d;
    aqs[i] = a.q_pool[(size_t)bh * Lq + i] * a.scale;
    lses[i] = a.lse[(size_t)bh * Lq + i];
  }
  __syncthreads();
  // phase 1: query side
  for (int i = threadIdx.x; i < Lq; i += kAtaThreads) {
    const float aq = aqs[i], ls = lses[i], d = Ds[i];
    float g[DVP];
#pragma unroll
    for (int e = 0; e < DVP; ++e) g[e] = gs[i * DVP + e];
    float acc = 0.f;
    for (int j = 0; j < Lk; ++j) {
      const float kj = ks[j];
      const float p = __expf(aq * kj - ls);
      float gp = 0.f;
#pragma unroll
      for (int e = 0; e < DVP; ++e) gp = fmaf(g[e], vs[j * DVP + e], gp);
      acc = fmaf(p * (gp - d), kj, acc);
    }
    scatter_group(a.g_qp + group_offset(((size_t)bh * Lq + i) * G, a.nf, a.blk_q, a.B), G, a.q_arg[(size_t)bh * Lq + i],
                  acc * a.scale);
  }
  // phase 2: key / value side
  for (int j = threadIdx.x; j < Lk; j += kAtaThreads) {
    const float kj = ks[j];
    float v[DVP], gv[DVP];
#pragma unroll
    for (int e = 0; e < DVP; ++e) { v[e] = vs[j * DVP + e]; gv[e] = 0.f; }
    float gk = 0.f;
    for (int i = 0; i < Lq; ++i) {
      const float aq = aqs[i];
      const float p = __expf(aq * kj - lses[i]);
      float gp = 0.f;
#pragma unroll
      for (int e = 0; e < DVP; ++e) {
        const float g = gs[i * DVP + e];
        gp = fmaf(g, v[e], gp);
        gv[e] = fmaf(p, g, gv[e]);
      }
      gk = fmaf(p * (gp - Ds[i]), aq, gk);
    }
    scatter_group(a.g_kp + group_offset(((size_t)bh * Lk + j) * G, a.nf, a.blk_k, a.B), G, a.k_arg[(size_t)bh * Lk + j], gk);
    float* o = a.g_v + (((size_t)b * Lk + j) * a.H + h) * DV;
#pragma unroll
    for (int e = 0; e < DVP; ++e)
      if (e < DV) o[e] = gv[e];
  }
}

size_t fwd_smem(int Lk, int dvp) { return ((size_t)round_up(Lk, 4) + (size_t)Lk * dvp) * sizeof(float); }
size_t bwd_smem(int Lq, int Lk, int dvp) {
  return fwd_smem(Lk, dvp) + ((size_t)3 * round_up(Lq, 4) + (size_t)Lq * dvp) * sizeof(float);
}

int dv_padded(int dv) { return dv <= 4 ? 4 : dv <= 8 ? 8 : dv <= 16 ? 16 : dv <= 32 ? 32 : 64; }

template <int DVP>
int launch(bool bwd, const AtaArgs& a, cudaStream_t st) {
  const size_t smem = bwd ? bwd_smem(a.Lq, a.Lk, DVP) : fwd_smem(a.Lk, DVP);
  if (smem > 200 * 1024) return GPBLUR_EUNSUPPORTED;
  const void* f = bwd ? (const void*)ata_bwd_kernel<DVP> : (const void*)ata_fwd_kernel<DVP>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   // per device
  ProfScope ps(ST_OTHER, st);
  const int grid = a.B * a.H;
  if (bwd) ata_bwd_kernel<DVP><<<grid, kAtaThreads, smem, st>>>(a);
  else ata_fwd_kernel<DVP><<<grid, kAtaThreads, smem, st>>>(a);
  note_launch();
  return check_launch(bwd ? "ata_backward" : "ata_forward");
}

int dispatch(bool bwd, const AtaArgs& a, cudaStream_t st) {
  switch (dv_padded(a.DV)) {
    case 4: return launch<4>(bwd, a, st);
    case 8: return launch<8>(bwd, a, st);
    case 16: return launch<16>(bwd, a, st);
    case 32: return launch<32>(bwd, a, st);
    default: return launch<64>(bwd, a, st);
  }
}

bool bad_dims(int B, int H, int Lq, int Lk, int G, int DV) {
  return B < 0 || H < 1 || Lq < 1 || Lk < 1 || G < 1 || DV < 1 || DV > 64;
}

}  // namespace

}  // namespace gpblur

using namespace gpblur;

static bool bad_layout(int nf, int B, int H, int Lq, int Lk, int G) {
  if (nf == 0) return false;
  if (nf < 0 || G % nf != 0) return true;
  const long long C = (long long)H * (G / nf);                  // channels of one filter stack = h * d_k
  return (C * Lq) % G != 0 || (C * Lk) % G != 0;
}

extern "C" int gpblur_ata_forward_fused_stacks(const float* qp, const float* kp, const float* v, long long v_sb,
                                               long long v_sh, long long v_sl, int B, int H, int Lq, int Lk, int G,
                                               int DV, int nf, float scale, float* ctx, float* q_pool, float* k_pool,
                                               int* q_arg, int* k_arg, float* lse, void* stream) {
  if (bad_dims(B, H, Lq, Lk, G, DV) || bad_layout(nf, B, H, Lq, Lk, G)) return GPBLUR_EINVAL;
  if (B == 0) return GPBLUR_OK;
  if (!qp || !kp || !v || !ctx || !q_pool || !k_pool || !q_arg || !k_arg || !lse) return GPBLUR_EINVAL;
  if ((G & 3) == 0 && (((uintptr_t)qp | (uintptr_t)kp) & 15)) return GPBLUR_EINVAL;
  AtaArgs a{};
  a.qp = qp; a.kp = kp; a.v = v; a.v_sb = v_sb; a.v_sh = v_sh; a.v_sl = v_sl;
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.G = G; a.DV = DV; a.scale = scale;
  a.nf = nf;
  if (nf) { a.blk_q = (long long)H * (G / nf) * Lq; a.blk_k = (long long)H * (G / nf) * Lk; }
  a.ctx = ctx; a.q_pool = q_pool; a.k_pool = k_pool; a.q_arg = q_arg; a.k_arg = k_arg; a.lse = lse;
  return dispatch(false, a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gpblur_ata_backward_fused_stacks(const float* g_ctx, const float* ctx, const float* v, long long v_sb,
                                                long long v_sh, long long v_sl, const float* q_pool,
                                                const float* k_pool, const int* q_arg, const int* k_arg,
                                                const float* lse, int B, int H, int Lq, int Lk, int G, int DV, int nf,
                                                float scale, float* g_qp, float* g_kp, float* g_v, void* stream) {
  if (bad_dims(B, H, Lq, Lk, G, DV) || bad_layout(nf, B, H, Lq, Lk, G)) return GPBLUR_EINVAL;
  if (B == 0) return GPBLUR_OK;
  if (!g_ctx || !ctx || !v || !q_pool || !k_pool || !q_arg || !k_arg || !lse || !g_qp || !g_kp || !g_v) return GPBLUR_EINVAL;
  if ((G & 3) == 0 && (((uintptr_t)g_qp | (uintptr_t)g_kp) & 15)) return GPBLUR_EINVAL;
  AtaArgs a{};
  a.v = v; a.v_sb = v_sb; a.v_sh = v_sh; a.v_sl = v_sl;
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.G = G; a.DV = DV; a.scale = scale;
  a.nf = nf;
  if (nf) { a.blk_q = (long long)H * (G / nf) * Lq; a.blk_k = (long long)H * (G / nf) * Lk; }
  a.ctx = const_cast<float*>(ctx); a.q_pool = const_cast<float*>(q_pool); a.k_pool = const_cast<float*>(k_pool);
  a.q_arg = const_cast<int*>(q_arg); a.k_arg = const_cast<int*>(k_arg); a.lse = const_cast<float*>(lse);
  a.g_ctx = g_ctx; a.g_qp = g_qp; a.g_kp = g_kp; a.g_v = g_v;
  return dispatch(true, a, reinterpret_cast<cudaStream_t>(stream));
}

extern "C" int gpblur_ata_forward(const float* qp, const float* kp, const float* v, long long v_sb, long long v_sh,
                                  long long v_sl, int B, int H, int Lq, int Lk, int G, int DV, float scale, float* ctx,
                                  float* q_pool, float* k_pool, int* q_arg, int* k_arg, float* lse, void* stream) {
  return gpblur_ata_forward_fused_stacks(qp, kp, v, v_sb, v_sh, v_sl, B, H, Lq, Lk, G, DV, 0, scale, ctx, q_pool, k_pool,
                                         q_arg, k_arg, lse, stream);
}

extern "C" int gpblur_ata_backward(const float* g_ctx, const float* ctx, const float* v, long long v_sb, long long v_sh,
                                   long long v_sl, const float* q_pool, const float* k_pool, const int* q_arg,
                                   const int* k_arg, const float* lse, int B, int H, int Lq, int Lk, int G, int DV,
                                   float scale, float* g_qp, float* g_kp, float* g_v, void* stream) {
  return gpblur_ata_backward_fused_stacks(g_ctx, ctx, v, v_sb, v_sh, v_sl, q_pool, k_pool, q_arg, k_arg, lse, B, H, Lq, Lk,
                                          G, DV, 0, scale, g_qp, g_kp, g_v, stream);
}
