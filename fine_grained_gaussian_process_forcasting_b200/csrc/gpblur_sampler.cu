// Window sampler on the device: the reference materialises every sampled window on the
// host (/root/reference/Utils/base_train.py:67-97: one pandas .iloc slice per window into float64 arrays of
// [max_samples, time_steps, F], then FloatTensor copies) and ships each batch with .to(device)
// (/root/reference/train.py:160-161).  Here the series table lives in HBM ONCE ([rows, F] fp32, entities back to back
// in time order) and a batch is gathered by one launch from the window start rows:
//
//   enc[b, t, :] = table[start_b + t, :]                          t < n_enc
//   dec[b, t, :] = table[start_b + n_enc + t, :]                  t < T - n_enc - pred_len
//   y[b, t]      = target[start_b + T - pred_len + t]             t < pred_len
//
// start_b < 0 marks a window the reference leaves zero-filled (max_samples larger than the number of valid sampling
// locations, base_train.py:56-69).  A window's rows are CONTIGUOUS in the table, so each output segment is a straight
// copy: pure byte movement, HBM / L2 bound, bit-exact.  One CTA walks windows b = blockIdx.x, + gridDim.x, ...; its
// threads stride over the window's floats (coalesced on both sides; 16-byte accesses when F is a multiple of 4, which
// keeps every segment 16-byte aligned).
#include "gpblur_common.cuh"

namespace gpblur {

namespace {

constexpr int kGatherThreads = 256;

struct GatherArgs {
  const float* table;
  const float* target;
  const long long* starts;
  long long rows, B;
  int F, T, n_enc, n_dec, pred_len;
  float* enc;
  float* dec;
  float* y;
};

template <typename V>   // V = float4 (F % 4 == 0, 16-byte aligned bases) or float
__global__ void __launch_bounds__(kGatherThreads) window_gather_kernel(GatherArgs a) {
  constexpr int VW = sizeof(V) / 4;
  const long long enc_n = (long long)a.n_enc * a.F / VW, dec_n = (long long)a.n_dec * a.F / VW;
  for (long long b = blockIdx.x; b < a.B; b += gridDim.x) {
    const long long start = a.starts[b];
    const bool live = start >= 0 && start + a.T <= a.rows;
    const V* src = reinterpret_cast<const V*>(a.table + (live ? start : 0) * a.F);
    V* e = reinterpret_cast<V*>(a.enc + b * a.n_enc * a.F);
    V* d = reinterpret_cast<V*>(a.dec + b * a.n_dec * a.F);
    V zero;
    memset(&zero, 0, sizeof(V));
    for (long long i = threadIdx.x; i < enc_n + dec_n; i += kGatherThreads) {
      const V v = live ? src[i] : zero;
      if (i < enc_n) e[i] = v;
      else d[i - enc_n] = v;
    }
    const float* ty = a.target + (live ? start + a.T - a.pred_len : 0);
    for (int t = threadIdx.x; t < a.pred_len; t += kGatherThreads) a.y[b * a.pred_len + t] = live ? ty[t] : 0.f;
  }
}

}  // namespace

}  // namespace gpblur

using namespace gpblur;

extern "C" int gpblur_window_gather(const float* table, const float* target, long long rows, int F,
                                    const long long* starts, long long B, int T, int n_enc, int pred_len, float* enc,
                                    float* dec, float* y, void* stream) {
  if (rows < 0 || B < 0 || F < 1 || T < 1 || n_enc < 0 || pred_len < 0 || n_enc + pred_len > T) return GPBLUR_EINVAL;
  const int n_dec = T - n_enc - pred_len;
  if (B == 0) return GPBLUR_OK;
  if (!starts || (rows > 0 && (!table || !target))) return GPBLUR_EINVAL;   // rows == 0: every window is dead
  if ((n_enc > 0 && !enc) || (n_dec > 0 && !dec) || (pred_len > 0 && !y)) return GPBLUR_EINVAL;
  GatherArgs args{table, target, starts, rows, B, F, T, n_enc, n_dec, pred_len, enc, dec, y};
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  // windows are a few KB each: several resident CTAs per SM keep enough loads in flight; at most one CTA per window
  long long grid = (long long)num_sms() * 8;
  if (grid > B) grid = B;
  const bool vec = (F % 4 == 0) && (((uintptr_t)table | (uintptr_t)enc | (uintptr_t)dec) % 16 == 0);
  ProfScope ps(ST_OTHER, st);
  if (vec) window_gather_kernel<float4><<<(int)grid, kGatherThreads, 0, st>>>(args);
  else window_gather_kernel<float><<<(int)grid, kGatherThreads, 0, st>>>(args);
  note_launch();
  return check_launch("window_gather");
}
