// The two steps on either side of the GP blur in one training step, each as ONE
// HBM-bound pass forward and ONE backward:
//
//   blur application   x_noisy = x + proj_up(mean)            /root/reference/denoising_model/denoise_model_2.py:32-40
//                      proj_up = nn.Linear(1, d): x_noisy[n, :] = x[n, :] + mean[n] * w_up + b_up
//   loss assembly      final = final_projection(h)            /root/reference/forecast_denoising.py:84
//                      mse   = mean((y - final)^2)            :103
//                      loss  = mse + clip(lam, 0, 0.005) * mll_error,  mll_error = -mean_b(elbo_b)     :87-89, :104
//
// Both are "row times vector + per-column sums" shapes: a row of D <= 128 floats is handled by TPR = D / 4 (rounded up
// to a power of two) neighbouring lanes with 16-byte accesses, row-wise dot products are reduced with shuffles inside
// those lanes, per-column sums are kept in registers across the rows of a CTA and combined CTA -> global partials ->
// the block that takes the last ticket (fixed order: bit-deterministic; self-resetting ticket word, see
// gpblur_elbo_backward_fused).  Reference semantics of torch.clip's gradient: passes where 0 <= lam <= 0.005.
#include "gpblur_common.cuh"

namespace gpblur {

namespace {

constexpr int kStepThreads = 256;
constexpr int kMaxParts = 1024;            // CTAs of the reducing kernels (size of the partial buffers)

__device__ __forceinline__ int tpr_of(int D) {           // lanes per row: power of two >= ceil(D / 4), <= 32
  int t = 1;
  while (t * 4 < D) t <<= 1;
  return t;
}

__device__ __forceinline__ float4 ld4(const float* p, int d, int D, bool vec) {
  if (vec) return *reinterpret_cast<const float4*>(p + d);
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (d < D) v.x = p[d];
  if (d + 1 < D) v.y = p[d + 1];
  if (d + 2 < D) v.z = p[d + 2];
  if (d + 3 < D) v.w = p[d + 3];
  return v;
}
__device__ __forceinline__ void st4(float* p, int d, int D, bool vec, float4 v) {
  if (vec) { *reinterpret_cast<float4*>(p + d) = v; return; }
  if (d < D) p[d] = v.x;
  if (d + 1 < D) p[d + 1] = v.y;
  if (d + 2 < D) p[d + 2] = v.z;
  if (d + 3 < D) p[d + 3] = v.w;
}
// sum over the TPR lanes of a row group (TPR a power of two <= 32, groups aligned)
__device__ __forceinline__ float group_sum(float v, int tpr) {
  for (int o = tpr >> 1; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kStepThreads) blur_apply_fwd_kernel(const float* __restrict__ x,
                                                                      const float* __restrict__ mean,
                                                                      const float* __restrict__ w_up,
                                                                      const float* __restrict__ b_up, long long N, int D,
                                                                      float* __restrict__ out) {
  const int tpr = tpr_of(D), rpi = kStepThreads / tpr;          // rows per iteration of a CTA
  const int sub = threadIdx.x % tpr, rloc = threadIdx.x / tpr, d = 4 * sub;
  const bool vec = (D & 3) == 0, on = d < D;
  const float4 w = on ? ld4(w_up, d, D, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float4 b = on ? ld4(b_up, d, D, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
  if (!on) return;
  // four rows per thread and iteration: all their loads are issued before the first store (a single 16-byte load in
  // flight per thread left the kernel latency-bound at ~0.7 TB/s)
  constexpr int U = 8;
  const long long stride = (long long)gridDim.x * rpi;
  for (long long n0 = (long long)blockIdx.x * rpi + rloc; n0 < N; n0 += U * stride) {
    float m[U];
    float4 xv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long n = n0 + u * stride;
      if (n < N) { m[u] = mean[n]; xv[u] = ld4(x + n * D, d, D, vec); }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long n = n0 + u * stride;
      if (n < N)
        st4(out + n * D, d, D, vec,
            make_float4(fmaf(m[u], w.x, xv[u].x + b.x), fmaf(m[u], w.y, xv[u].y + b.y), fmaf(m[u], w.z, xv[u].z + b.z),
                        fmaf(m[u], w.w, xv[u].w + b.w)));
    }
  }
}

// Sum of one double per thread over the CTA in a FIXED tree order (bit-deterministic); valid in every thread.
__device__ __forceinline__ double block_sum_fixed(double v) {
  __shared__ double bs[kStepThreads];
  __syncthreads();                       // a previous use of bs is over
  bs[threadIdx.x] = v;
  __syncthreads();
  for (int o = kStepThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) bs[threadIdx.x] += bs[threadIdx.x + o];
    __syncthreads();
  }
  return bs[0];
}

// Shared tail of the reducing kernels: per-thread column partials (two float4 per thread: `a` and `b` sums of its 4
// columns) -> CTA partial -> global partial [gridDim.x][2][Dp] -> the last block sums all of them in block order.
// Returns true in the block that holds the totals (in smem tot[2][128], valid for all its threads after the call).
__device__ __forceinline__ bool column_totals(float4 sa, float4 sb, int tpr, int D, float* __restrict__ partial,
                                              unsigned* __restrict__ ticket, float (*tot)[128]) {
  __shared__ float red[2][kStepThreads / 1][4];                 // [a|b][thread][4]
  const int sub = threadIdx.x % tpr, rloc = threadIdx.x / tpr, rpi = kStepThreads / tpr;
  red[0][threadIdx.x][0] = sa.x; red[0][threadIdx.x][1] = sa.y; red[0][threadIdx.x][2] = sa.z; red[0][threadIdx.x][3] = sa.w;
  red[1][threadIdx.x][0] = sb.x; red[1][threadIdx.x][1] = sb.y; red[1][threadIdx.x][2] = sb.z; red[1][threadIdx.x][3] = sb.w;
  __syncthreads();
  const int Dp = 4 * tpr;
  if (rloc == 0) {
    // thread `sub` of the first row group adds the rpi row groups in order
#pragma unroll
    for (int which = 0; which < 2; ++which) {
      float t[4] = {0.f, 0.f, 0.f, 0.f};
      for (int r = 0; r < rpi; ++r)
#pragma unroll
        for (int e = 0; e < 4; ++e) t[e] += red[which][r * tpr + sub][e];
#pragma unroll
      for (int e = 0; e < 4; ++e) partial[((size_t)blockIdx.x * 2 + which) * Dp + 4 * sub + e] = t[e];
    }
  }
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return false;
  __threadfence();
  for (int i = threadIdx.x; i < 2 * Dp; i += kStepThreads) {
    const int which = i / Dp, c = i - which * Dp;
    // four independent chains over g = 0, 4, 8, ... | 1, 5, ... | ... (fixed order), L2 loads (__ldcg: the partials were
    // written by other CTAs and fenced), so that many loads are in flight: one serial chain of ~600 loads cost 20 us
    const float* src = partial + (size_t)which * Dp + c;
    const size_t gs = (size_t)2 * Dp;
    double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
    unsigned g = 0;
#pragma unroll 4
    for (; g + 4 <= gridDim.x; g += 4) {
      t0 += (double)__ldcg(src + (size_t)g * gs);
      t1 += (double)__ldcg(src + (size_t)(g + 1) * gs);
      t2 += (double)__ldcg(src + (size_t)(g + 2) * gs);
      t3 += (double)__ldcg(src + (size_t)(g + 3) * gs);
    }
    for (; g < gridDim.x; ++g) t0 += (double)__ldcg(src + (size_t)g * gs);
    tot[which][c] = (float)((t0 + t1) + (t2 + t3));
  }
  __syncthreads();
  (void)D;
  return true;
}

// g_out [N, D] -> g_mean [N] = g_out . w_up,  g_w [D] = sum_n mean[n] g_out[n, :],  g_b [D] = sum_n g_out[n, :]
// (the gradient of x is g_out itself: no kernel)
__global__ void __launch_bounds__(kStepThreads) blur_apply_bwd_kernel(const float* __restrict__ g_out,
                                                                      const float* __restrict__ mean,
                                                                      const float* __restrict__ w_up, long long N, int D,
                                                                      float* __restrict__ g_mean, float* __restrict__ g_w,
                                                                      float* __restrict__ g_b, float* __restrict__ partial,
                                                                      unsigned* __restrict__ ticket) {
  const int tpr = tpr_of(D), rpi = kStepThreads / tpr;
  const int sub = threadIdx.x % tpr, rloc = threadIdx.x / tpr, d = 4 * sub;
  const bool vec = (D & 3) == 0, on = d < D;
  const float4 w = on ? ld4(w_up, d, D, vec) : make_float4(0.f, 0.f, 0.f, 0.f);
  float4 sw = make_float4(0.f, 0.f, 0.f, 0.f), sb = sw;
  const long long n_iter = (N + rpi - 1) / rpi;
  constexpr int U = 8;                                                 // loads of eight row groups in flight per thread
  for (long long it0 = blockIdx.x; it0 < n_iter; it0 += (long long)U * gridDim.x) {   // whole row groups: shuffles stay convergent
    float4 gu[U];
    float mu[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long n = (it0 + (long long)u * gridDim.x) * rpi + rloc;
      gu[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      mu[u] = 0.f;
      if (n < N && on) { gu[u] = ld4(g_out + n * D, d, D, vec); mu[u] = mean[n]; }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long n = (it0 + (long long)u * gridDim.x) * rpi + rloc;
      const float4 g = gu[u];
      const float m = mu[u];
      float dot = g.x * w.x + g.y * w.y + g.z * w.z + g.w * w.w;
      dot = group_sum(dot, tpr);
      if (n < N && sub == 0) g_mean[n] = dot;
      sw.x = fmaf(m, g.x, sw.x); sw.y = fmaf(m, g.y, sw.y); sw.z = fmaf(m, g.z, sw.z); sw.w = fmaf(m, g.w, sw.w);
      sb.x += g.x; sb.y += g.y; sb.z += g.z; sb.w += g.w;
    }
  }
  __shared__ float tot[2][128];
  if (!column_totals(sw, sb, tpr, D, partial, ticket, tot)) return;
  for (int c = threadIdx.x; c < D; c += kStepThreads) {
    g_w[c] = tot[0][c];
    g_b[c] = tot[1][c];
  }
}

// offset of row n = (b, p) of a [B', P, D] view with batch stride h_bstride.  The 64-bit division cost ~150 instructions
// per row (more than the rest of the row's work): rows are counted in 32 bits whenever they fit, and a contiguous view
// needs no division at all.
__device__ __forceinline__ long long row_offset(long long n, int P, long long h_bstride, int D, bool small) {
  if (h_bstride == (long long)P * D) return n * D;
  if (small) {
    const unsigned q = (unsigned)n / (unsigned)P, r = (unsigned)n - q * (unsigned)P;
    return (long long)q * h_bstride + (long long)r * D;
  }
  return (n / P) * h_bstride + (n % P) * (long long)D;
}

// ---------------------------------------------------------------------------------------------------------------
// final[n] = h[n, :] . w_f + b_f ; partial sums of (y - final)^2 ; the last block: mse, mll_error, loss
__global__ void __launch_bounds__(kStepThreads) loss_fwd_kernel(const float* __restrict__ h, long long h_bstride, int P,
                                                                const float* __restrict__ w_f, const float* __restrict__ b_f,
                                                                const float* __restrict__ y, const float* __restrict__ elbo,
                                                                long long B, const float* __restrict__ lam, long long N,
                                                                int D, float* __restrict__ final_out,
                                                                float* __restrict__ scalars, float* __restrict__ partial,
                                                                unsigned* __restrict__ ticket) {
  const int tpr = tpr_of(D), rpi = kStepThreads / tpr;
  const int sub = threadIdx.x % tpr, rloc = threadIdx.x / tpr, d = 4 * sub;
  const bool vec = (D & 3) == 0 && (h_bstride & 3) == 0, on = d < D;
  const float4 w = on ? ld4(w_f, d, D, (D & 3) == 0) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float bf = b_f[0];
  float se = 0.f;
  const long long n_iter = (N + rpi - 1) / rpi;
  constexpr int U = 8;                                                 // loads of eight row groups in flight per thread
  for (long long it0 = blockIdx.x; it0 < n_iter; it0 += (long long)U * gridDim.x) {
    float4 hu[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long n = (it0 + (long long)u * gridDim.x) * rpi + rloc;
      hu[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < N && on) hu[u] = ld4(h + row_offset(n, P, h_bstride, D, N < (1ll << 31)), d, D, vec);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long n = (it0 + (long long)u * gridDim.x) * rpi + rloc;
      const float4 hv = hu[u];
      float dot = hv.x * w.x + hv.y * w.y + hv.z * w.z + hv.w * w.w;
      dot = group_sum(dot, tpr);
      if (n < N && sub == 0) {
        const float f = dot + bf;
        final_out[n] = f;
        if (y) { const float e = y[n] - f; se = fmaf(e, e, se); }
      }
    }
  }
  // squared-error partial of the CTA (fixed order), then the last block finishes
  const double cta_se = block_sum_fixed((double)se);
  if (threadIdx.x == 0) partial[blockIdx.x] = (float)cta_se;
  __shared__ bool last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    last = atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  // the last block: all its threads sum strided subsets (fixed order), then the fixed tree
  double sse = 0.0;
  for (unsigned g = threadIdx.x; g < gridDim.x; g += kStepThreads) sse += (double)__ldcg(partial + g);
  sse = block_sum_fixed(sse);
  double es = 0.0;
  if (elbo)
    for (long long b = threadIdx.x; b < B; b += kStepThreads) es += (double)elbo[b];
  es = block_sum_fixed(es);
  if (threadIdx.x != 0) return;
  const double mse = y ? sse / (double)N : 0.0;
  const double mll_error = elbo ? -es / (double)B : 0.0;
  const float lm = lam ? lam[0] : 0.f;
  const double lc = lm < 0.f ? 0.0 : (lm > 0.005f ? 0.005 : (double)lm);
  scalars[0] = (float)(mse + lc * mll_error);     // loss
  scalars[1] = (float)mse;
  scalars[2] = (float)mll_error;
}

// upstream: g_final [N] (nullable), g_loss, g_mse (device scalars, nullable = 0).  d loss / d final = 2 (final - y) / N.
//   g_f[n]  = g_final[n] + (g_loss + g_mse) * 2 (final[n] - y[n]) / N
//   g_h     = g_f[n] * w_f ;  g_w_f = sum_n g_f[n] h[n, :] ;  g_b_f = sum_n g_f[n]
//   g_elbo_b = -g_loss * clip(lam) / B ;  g_lam = g_loss * mll_error * [0 <= lam <= 0.005]
__global__ void __launch_bounds__(kStepThreads) loss_bwd_kernel(const float* __restrict__ h, long long h_bstride, int P,
                                                                const float* __restrict__ w_f, const float* __restrict__ y,
                                                                const float* __restrict__ final_in,
                                                                const float* __restrict__ scalars,
                                                                const float* __restrict__ lam,
                                                                const float* __restrict__ g_final,
                                                                const float* __restrict__ g_loss,
                                                                const float* __restrict__ g_mse, long long B, long long N,
                                                                int D, float* __restrict__ g_h, float* __restrict__ g_w,
                                                                float* __restrict__ g_b, float* __restrict__ g_elbo,
                                                                float* __restrict__ g_lam, float* __restrict__ partial,
                                                                unsigned* __restrict__ ticket) {
  const int tpr = tpr_of(D), rpi = kStepThreads / tpr;
  const int sub = threadIdx.x % tpr, rloc = threadIdx.x / tpr, d = 4 * sub;
  const bool vecw = (D & 3) == 0, vec = vecw && (h_bstride & 3) == 0, on = d < D;
  const float4 w = on ? ld4(w_f, d, D, vecw) : make_float4(0.f, 0.f, 0.f, 0.f);
  const float gl = g_loss ? g_loss[0] : 0.f, gm = g_mse ? g_mse[0] : 0.f;
  const float c2 = y ? (gl + gm) * 2.0f / (float)N : 0.f;
  float4 sw = make_float4(0.f, 0.f, 0.f, 0.f), sb = sw;
  const long long n_iter = (N + rpi - 1) / rpi;
  constexpr int U = 8;                                                 // loads of eight row groups in flight per thread
  for (long long it0 = blockIdx.x; it0 < n_iter; it0 += (long long)U * gridDim.x) {
    float4 hu[U];
    float gfu[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long n = (it0 + (long long)u * gridDim.x) * rpi + rloc;
      hu[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      gfu[u] = 0.f;
      if (n < N && on) {
        float gf = g_final ? g_final[n] : 0.f;
        if (y) gf = fmaf(c2, final_in[n] - y[n], gf);
        gfu[u] = gf;
        hu[u] = ld4(h + row_offset(n, P, h_bstride, D, N < (1ll << 31)), d, D, vec);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long n = (it0 + (long long)u * gridDim.x) * rpi + rloc;
      if (n >= N || !on) continue;
      const float gf = gfu[u];
      const float4 hv = hu[u];
      if (g_h) st4(g_h + n * D, d, D, vecw, make_float4(gf * w.x, gf * w.y, gf * w.z, gf * w.w));
      sw.x = fmaf(gf, hv.x, sw.x); sw.y = fmaf(gf, hv.y, sw.y); sw.z = fmaf(gf, hv.z, sw.z); sw.w = fmaf(gf, hv.w, sw.w);
      if (sub == 0) sb.x += gf;
    }
  }
  // elbo / lam gradients: a few elements, block 0
  if (blockIdx.x == 0) {
    const float lm = lam ? lam[0] : 0.f;
    const float lc = lm < 0.f ? 0.f : (lm > 0.005f ? 0.005f : lm);
    if (g_elbo)
      for (long long b = threadIdx.x; b < B; b += kStepThreads) g_elbo[b] = -gl * lc / (float)B;
    if (g_lam && threadIdx.x == 0) g_lam[0] = (lm >= 0.f && lm <= 0.005f) ? gl * scalars[2] : 0.f;
  }
  __shared__ float tot[2][128];
  if (!column_totals(sw, sb, tpr, D, partial, ticket, tot)) return;
  for (int c = threadIdx.x; c < D; c += kStepThreads) g_w[c] = tot[0][c];
  if (threadIdx.x == 0) g_b[0] = tot[1][0];
}

int step_grid(long long N, int D) {
  int tpr = 1;
  while (tpr * 4 < D) tpr <<= 1;
  const int rpi = kStepThreads / tpr;
  long long g = (N + rpi - 1) / rpi;
  const long long cap = 2 * 148;                 // 2 CTAs per SM x 8 row groups in flight per thread (~10 MB of loads in
                                                 // flight); few CTAs keep the last block's pass over the partials short
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

}  // namespace

}  // namespace gpblur

using namespace gpblur;

extern "C" {

size_t gpblur_step_scratch_floats(int D) {
  (void)D;
  return (size_t)kMaxParts * 2 * 128;
}

int gpblur_blur_apply_forward(const float* x, const float* mean, const float* w_up, const float* b_up, long long N,
                              int D, float* out, void* stream) {
  if (N < 0 || D < 1 || D > 128) return GPBLUR_EINVAL;
  if (N == 0) return GPBLUR_OK;
  if (!x || !mean || !w_up || !b_up || !out) return GPBLUR_EINVAL;
  blur_apply_fwd_kernel<<<step_grid(N, D), kStepThreads, 0, (cudaStream_t)stream>>>(x, mean, w_up, b_up, N, D, out);
  note_launch();
  return check_launch("blur_apply_fwd");
}

int gpblur_blur_apply_backward(const float* g_out, const float* mean, const float* w_up, long long N, int D,
                               float* g_mean, float* g_w, float* g_b, float* scratch, unsigned* ticket, void* stream) {
  if (N < 1 || D < 1 || D > 128) return GPBLUR_EINVAL;
  if (!g_out || !mean || !w_up || !g_mean || !g_w || !g_b || !scratch || !ticket) return GPBLUR_EINVAL;
  blur_apply_bwd_kernel<<<step_grid(N, D), kStepThreads, 0, (cudaStream_t)stream>>>(g_out, mean, w_up, N, D, g_mean, g_w,
                                                                                      g_b, scratch, ticket);
  note_launch();
  return check_launch("blur_apply_bwd");
}

int gpblur_loss_forward(const float* h, long long h_bstride, int P, const float* w_f, const float* b_f, const float* y,
                        const float* elbo, long long B, const float* lam, long long N, int D, float* final_out,
                        float* scalars, float* scratch, unsigned* ticket, void* stream) {
  if (N < 1 || D < 1 || D > 128 || B < 0 || P < 1 || N % P != 0 || h_bstride < (long long)P * D) return GPBLUR_EINVAL;
  if (!h || !w_f || !b_f || !final_out || !scalars || !scratch || !ticket) return GPBLUR_EINVAL;
  loss_fwd_kernel<<<step_grid(N, D), kStepThreads, 0, (cudaStream_t)stream>>>(h, h_bstride, P, w_f, b_f, y, elbo, B, lam, N,
                                                                                D, final_out, scalars, scratch, ticket);
  note_launch();
  return check_launch("loss_fwd");
}

int gpblur_loss_backward(const float* h, long long h_bstride, int P, const float* w_f, const float* y, const float* final_in,
                         const float* scalars, const float* lam, const float* g_final, const float* g_loss,
                         const float* g_mse, long long B, long long N, int D, float* g_h, float* g_w, float* g_b,
                         float* g_elbo, float* g_lam, float* scratch, unsigned* ticket, void* stream) {
  if (N < 1 || D < 1 || D > 128 || B < 0 || P < 1 || N % P != 0 || h_bstride < (long long)P * D) return GPBLUR_EINVAL;
  if (!h || !w_f || !final_in || !scalars || !g_w || !g_b || !scratch || !ticket) return GPBLUR_EINVAL;
  loss_bwd_kernel<<<step_grid(N, D), kStepThreads, 0, (cudaStream_t)stream>>>(h, h_bstride, P, w_f, y, final_in, scalars, lam,
                                                                                g_final, g_loss, g_mse, B, N, D, g_h, g_w,
                                                                                g_b, g_elbo, g_lam, scratch, ticket);
  note_launch();
  return check_launch("loss_bwd");
}

}  // extern "C"
