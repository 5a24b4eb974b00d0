// N-reduction GEMMs of the backward:  C[p][q] = sum_n sc[n] * U[n][p] * V[n][q]   (split over N, deterministic)
//   Gram   : U = V = A (saved whitened cross-covariance), sc = g_var   -> S = sum_n g_var a a^T  (lower tile
//            triangle) and, on the diagonal tiles, u = sum_n g_mu a
//   W^T X  : U = W (kbar o k), V = X (raw inputs), sc = 1              -> feeds dZ and d lengthscale
// Both operands are K-major in memory ([n][cols]), i.e. already in the layout an FP32 FFMA register-tiled
// GEMM wants; partial results per split are summed in fixed order by the M x M backward stage.
#include "gpblur_common.cuh"

namespace gpblur {

namespace {

template <int R>
__device__ __forceinline__ int frag_col(int t, int r, int T) {
  if (R == 8) return r < 4 ? t * 4 + r : T / 2 + t * 4 + (r - 4);
  return t * R + r;
}

template <int R>
__device__ __forceinline__ void load_frag(float (&v)[R], const float* row, int t, int T) {
  if (R == 8) {
    const float4 a = *reinterpret_cast<const float4*>(row + t * 4);
    const float4 b = *reinterpret_cast<const float4*>(row + T / 2 + t * 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[R - 4] = b.x; v[R - 3] = b.y; v[R - 2] = b.z; v[R - 1] = b.w;
  } else if (R == 4) {
    const float4 a = *reinterpret_cast<const float4*>(row + t * 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
  } else if (R == 2) {
    const float2 a = *reinterpret_cast<const float2*>(row + t * 2);
    v[0] = a.x; v[1] = a.y;
  } else {
    v[0] = row[t];
  }
}

struct ReduceArgs {
  const float* U;      // [N][ldu]
  const float* V;      // [N][ldv]
  const float* sc;     // [N] or null (1)
  const float* gm;     // [N] (Gram only)
  float* C;            // [splits][P][ldc]
  float* uvec;         // [splits][P] (Gram only)
  long long N;
  int ldu, ldv, vcols; // vcols = valid columns of V (D for X)
  int P, ldc;
  int rows_per_split;
};

template <int TP, int TQ, bool GRAM>
__global__ void __launch_bounds__(kThreads, 2) reduce_gemm_kernel(ReduceArgs a) {
  constexpr int RP = TP / 16, RQ = TQ / 16;
  __shared__ __align__(16) float Us[2][kKS][TP];
  __shared__ __align__(16) float Vs[2][kKS][TQ];
  __shared__ float scs[2][kKS];
  __shared__ float gms[2][kKS];

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  int tp, tq;
  if (GRAM) {
    int t = blockIdx.x;
    int i = (int)((sqrtf(8.0f * (float)t + 1.0f) - 1.0f) * 0.5f);
    while ((i + 1) * (i + 2) / 2 <= t) ++i;
    while (i * (i + 1) / 2 > t) --i;
    tp = i;
    tq = t - i * (i + 1) / 2;
  } else {
    tp = blockIdx.x;
    tq = 0;
  }
  const int p0 = tp * TP, q0 = tq * TQ;
  const long long r0 = (long long)blockIdx.y * a.rows_per_split;
  long long r1 = r0 + a.rows_per_split;
  if (r1 > a.N) r1 = a.N;
  const bool diag = GRAM && (tp == tq);
  const bool vvec = GRAM || ((a.ldv % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.V) & 15) == 0));

  float acc[RP][RQ];
#pragma unroll
  for (int i = 0; i < RP; ++i)
#pragma unroll
    for (int j = 0; j < RQ; ++j) acc[i][j] = 0.f;
  float uacc[RP];
#pragma unroll
  for (int i = 0; i < RP; ++i) uacc[i] = 0.f;

  auto stage = [&](int buf, long long row0) {
    // U slice: [KS][TP] via cp.async (rows clamped; their scale is forced to 0)
    for (int idx = tid; idx < kKS * (TP / 4); idx += kThreads) {
      const int r = idx / (TP / 4), cq = idx - r * (TP / 4);
      long long n = row0 + r;
      if (n >= a.N) n = a.N - 1;
      cp_async16(&Us[buf][r][cq * 4], a.U + (size_t)n * a.ldu + p0 + cq * 4);
    }
    if (vvec) {
      for (int idx = tid; idx < kKS * (TQ / 4); idx += kThreads) {
        const int r = idx / (TQ / 4), cq = idx - r * (TQ / 4);
        long long n = row0 + r;
        if (n >= a.N) n = a.N - 1;
        if (q0 + cq * 4 < a.vcols) cp_async16(&Vs[buf][r][cq * 4], a.V + (size_t)n * a.ldv + q0 + cq * 4);
        else *reinterpret_cast<float4*>(&Vs[buf][r][cq * 4]) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    } else {
      for (int idx = tid; idx < kKS * TQ; idx += kThreads) {
        const int r = idx / TQ, c = idx - r * TQ;
        long long n = row0 + r;
        if (n >= a.N) n = a.N - 1;
        Vs[buf][r][c] = (q0 + c < a.vcols) ? a.V[(size_t)n * a.ldv + q0 + c] : 0.f;
      }
    }
    if (tid < kKS) {
      const long long n = row0 + tid;
      const bool ok = n < r1;
      scs[buf][tid] = ok ? (a.sc ? a.sc[n] : 1.f) : 0.f;
      gms[buf][tid] = (ok && GRAM) ? a.gm[n] : 0.f;
    }
    cp_async_commit();
  };

  const int nsl = (int)((r1 - r0 + kKS - 1) / kKS);
  if (nsl > 0) stage(0, r0);
  for (int s = 0; s < nsl; ++s) {
    if (s + 1 < nsl) {
      stage((s + 1) & 1, r0 + (long long)(s + 1) * kKS);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const int buf = s & 1;
#pragma unroll
    for (int k = 0; k < kKS; ++k) {
      float av[RP], bv[RQ];
      load_frag<RP>(av, &Us[buf][k][0], ty, TP);
      load_frag<RQ>(bv, &Vs[buf][k][0], tx, TQ);
      const float sc = scs[buf][k];
      if (GRAM) {
        const float g = gms[buf][k];
#pragma unroll
        for (int i = 0; i < RP; ++i) uacc[i] = fmaf(g, av[i], uacc[i]);
      }
#pragma unroll
      for (int i = 0; i < RP; ++i) {
        const float as = av[i] * sc;
#pragma unroll
        for (int j = 0; j < RQ; ++j) acc[i][j] = fmaf(as, bv[j], acc[i][j]);
      }
    }
    __syncthreads();
  }

  float* Cs = a.C + (size_t)blockIdx.y * a.P * a.ldc;
#pragma unroll
  for (int i = 0; i < RP; ++i) {
    const int p = p0 + frag_col<RP>(ty, i, TP);
#pragma unroll
    for (int j = 0; j < RQ; ++j) {
      const int q = q0 + frag_col<RQ>(tx, j, TQ);
      Cs[(size_t)p * a.ldc + q] = acc[i][j];
    }
    if (diag && tx == 0) a.uvec[(size_t)blockIdx.y * a.P + p] = uacc[i];
  }
}

template <int TP, int TQ, bool GRAM>
int launch_reduce(const ReduceArgs& a, int ntiles, int splits, cudaStream_t st) {
  dim3 grid(ntiles, splits);
  ProfScope ps(GRAM ? ST_GRAM : ST_WX, st);
  reduce_gemm_kernel<TP, TQ, GRAM><<<grid, kThreads, 0, st>>>(a);
  note_launch();
  return check_launch("reduce_gemm");
}

template <int TP>
int launch_wx(const ReduceArgs& a, int DP, int ntiles, int splits, cudaStream_t st) {
  switch (DP) {
    case 16: return launch_reduce<TP, 16, false>(a, ntiles, splits, st);
    case 32: return launch_reduce<TP, 32, false>(a, ntiles, splits, st);
    case 64: return launch_reduce<TP, 64, false>(a, ntiles, splits, st);
    default: return launch_reduce<TP, 128, false>(a, ntiles, splits, st);
  }
}

}  // namespace

bool tc_reductions_supported(const WsLayout& L);
int launch_tc_reductions(const WsLayout& L, void* ws, const float* x, cudaStream_t st);

int launch_reductions(const WsLayout& L, void* ws, const float* x, cudaStream_t st) {
  if (L.N <= 0) return GPBLUR_OK;
  if (tc_reductions_supported(L)) return launch_tc_reductions(L, ws, x, st);   // tcgen05 3xTF32 path
  const int MP = L.MP;
  const int tp = MP < 128 ? MP : 128;
  const int nt = MP / tp;
  const float* gsc = ws_cptr<float>(ws, L.gsc);
  auto rows = [&](int splits) {
    long long r = (L.N + splits - 1) / splits;
    return (int)round_up_ll(r, kKS);
  };
  // Gram
  ReduceArgs g{};
  g.U = ws_cptr<float>(ws, L.A); g.V = g.U; g.sc = gsc + L.N; g.gm = gsc;
  g.C = ws_ptr<float>(ws, L.Spart); g.uvec = ws_ptr<float>(ws, L.upart);
  g.N = L.N; g.ldu = MP; g.ldv = MP; g.vcols = MP; g.P = MP; g.ldc = MP;
  g.rows_per_split = rows(L.splitsS);
  int rc;
  if (tp == 32) rc = launch_reduce<32, 32, true>(g, nt * (nt + 1) / 2, L.splitsS, st);
  else if (tp == 64) rc = launch_reduce<64, 64, true>(g, nt * (nt + 1) / 2, L.splitsS, st);
  else rc = launch_reduce<128, 128, true>(g, nt * (nt + 1) / 2, L.splitsS, st);
  if (rc) return rc;
  // W^T X
  ReduceArgs w{};
  w.U = ws_cptr<float>(ws, L.W); w.V = x; w.sc = nullptr; w.gm = nullptr;
  w.C = ws_ptr<float>(ws, L.WXpart); w.uvec = nullptr;
  w.N = L.N; w.ldu = MP; w.ldv = L.D; w.vcols = L.D; w.P = MP; w.ldc = L.DP;
  w.rows_per_split = rows(L.splitsZ);
  if (tp == 32) return launch_wx<32>(w, L.DP, nt, L.splitsZ, st);
  if (tp == 64) return launch_wx<64>(w, L.DP, nt, L.splitsZ, st);
  return launch_wx<128>(w, L.DP, nt, L.splitsZ, st);
}

}  // namespace gpblur
