// Backward of the dense ScaleKernel(RBF) covariance (gpblur_rbf_covariance): what torch autograd computes through
// gpytorch's kernel evaluation when an exact GP is trained (/root/reference/denoising_model/GPModel.py:4-13 builds
// MultivariateNormal(mean, ScaleKernel(RBF)(x)); the marginal log likelihood differentiates the covariance w.r.t. the
// raw lengthscale / outputscale and the inputs).  With W = G o K (G = upstream gradient of the covariance):
//   g_x1[i, d] = -sum_j W_ij (x1_id - x2_jd) / l_d^2          g_x2[j, d] = +sum_i W_ij (x1_id - x2_jd) / l_d^2
//   g_l_d      =  sum_ij W_ij (x1_id - x2_jd)^2 / l_d^3        g_os = sum_ij W_ij / os
// and the chain rule through softplus (sigmoid of the raw value); an isotropic kernel sums g_l over d.
// Three launches: W (tile kernel, as the forward) -> row pass (g_x1, per-row lengthscale / W sums) and column pass
// (g_x2) -> one CTA adds the per-row sums in row order (bit-deterministic).  O(n1 n2 D) work, exact-GP sizes.
#include "gpblur_common.cuh"

namespace gpblur {

namespace {

__device__ __forceinline__ float softplus_f(float r) { return r > 20.f ? r : log1pf(expf(r)); }
__device__ __forceinline__ float sigmoid_f(float r) { return 1.0f / (1.0f + expf(-r)); }

// W[i, j] = G[i, j] * os * exp(-1/2 sum_d ((x1_id - x2_jd) / l_d)^2): 16 x 16 tile per CTA, direct differences
__global__ void rbf_w_kernel(const float* __restrict__ x1, const float* __restrict__ x2, long long n1, long long n2, int D,
                             const float* __restrict__ raw_ell, int ard, const float* __restrict__ raw_os,
                             const float* __restrict__ g_out, float* __restrict__ W) {
  __shared__ float a[16][33], b[16][33], ie[32];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const long long i0 = (long long)blockIdx.y * 16, j0 = (long long)blockIdx.x * 16;
  float acc = 0.f;
  for (int d0 = 0; d0 < D; d0 += 32) {
    __syncthreads();
    if (threadIdx.x < 32) {
      const int d = d0 + threadIdx.x;
      ie[threadIdx.x] = d < D ? 1.0f / softplus_f(raw_ell[ard ? d : 0]) : 0.f;
    }
    for (int idx = threadIdx.x; idx < 16 * 32; idx += blockDim.x) {
      const int r = idx >> 5, c = idx & 31, d = d0 + c;
      a[r][c] = (i0 + r < n1 && d < D) ? x1[(size_t)(i0 + r) * D + d] : 0.f;
      b[r][c] = (j0 + r < n2 && d < D) ? x2[(size_t)(j0 + r) * D + d] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      const float df = (a[ty][c] - b[tx][c]) * ie[c];
      acc = fmaf(df, df, acc);
    }
  }
  const long long i = i0 + ty, j = j0 + tx;
  if (i < n1 && j < n2) W[(size_t)i * n2 + j] = g_out[(size_t)i * n2 + j] * softplus_f(raw_os[0]) * expf(-0.5f * acc);
}

// one CTA per row i of x1: threads = dimensions (d < D <= 128), loop over j.
//   g_x1[i, d] ; rowstat[i, d] = sum_j W_ij (x1_id - x2_jd)^2 ; rowstat[i, D] = sum_j W_ij   (double)
__global__ void rbf_row_kernel(const float* __restrict__ x1, const float* __restrict__ x2, long long n2, int D,
                               const float* __restrict__ raw_ell, int ard, const float* __restrict__ W,
                               float* __restrict__ g_x1, double* __restrict__ rowstat) {
  const long long i = blockIdx.x;
  const int d = threadIdx.x;
  const float* wr = W + (size_t)i * n2;
  if (d < D) {
    const float ell = softplus_f(raw_ell[ard ? d : 0]);
    const float xi = x1[(size_t)i * D + d];
    double s1 = 0.0, s2 = 0.0;
    for (long long j = 0; j < n2; ++j) {
      const float w = wr[j], df = xi - x2[(size_t)j * D + d];
      s1 += (double)(w * df);
      s2 += (double)(w * df * df);
    }
    if (g_x1) g_x1[(size_t)i * D + d] = (float)(-s1 / ((double)ell * ell));
    rowstat[(size_t)i * (D + 1) + d] = s2;
  } else if (d == D) {
    double s = 0.0;
    for (long long j = 0; j < n2; ++j) s += (double)wr[j];
    rowstat[(size_t)i * (D + 1) + D] = s;
  }
}

// one CTA per column j: g_x2[j, d] = + sum_i W_ij (x1_id - x2_jd) / l_d^2
__global__ void rbf_col_kernel(const float* __restrict__ x1, const float* __restrict__ x2, long long n1, long long n2,
                               int D, const float* __restrict__ raw_ell, int ard, const float* __restrict__ W,
                               float* __restrict__ g_x2) {
  const long long j = blockIdx.x;
  const int d = threadIdx.x;
  if (d >= D) return;
  const float ell = softplus_f(raw_ell[ard ? d : 0]);
  const float xj = x2[(size_t)j * D + d];
  double s1 = 0.0;
  for (long long i = 0; i < n1; ++i) s1 += (double)(W[(size_t)i * n2 + j] * (x1[(size_t)i * D + d] - xj));
  g_x2[(size_t)j * D + d] = (float)(s1 / ((double)ell * ell));
}

// rows added in row order: g_raw_ell [D] (ard) or [1], g_raw_os [1]
__global__ void rbf_finish_kernel(const double* __restrict__ rowstat, long long n1, int D, const float* __restrict__ raw_ell,
                                  int ard, const float* __restrict__ raw_os, float* __restrict__ g_raw_ell,
                                  float* __restrict__ g_raw_os) {
  __shared__ double iso[129];
  const int d = threadIdx.x;
  double t = 0.0;
  if (d <= D)
    for (long long i = 0; i < n1; ++i) t += rowstat[(size_t)i * (D + 1) + d];
  if (d < D) {
    const float r = raw_ell[ard ? d : 0];
    const double ell = (double)softplus_f(r);
    const double gl = t / (ell * ell * ell) * (double)sigmoid_f(r);
    if (ard) g_raw_ell[d] = (float)gl;
    iso[d] = gl;
  } else if (d == D) {
    const float r = raw_os[0];
    g_raw_os[0] = (float)(t / (double)softplus_f(r) * (double)sigmoid_f(r));
  }
  __syncthreads();
  if (!ard && d == 0) {
    double s = 0.0;
    for (int k = 0; k < D; ++k) s += iso[k];
    g_raw_ell[0] = (float)s;
  }
}

}  // namespace

}  // namespace gpblur

using namespace gpblur;

extern "C" size_t gpblur_rbf_covariance_backward_scratch_bytes(long long n1, long long n2, int D) {
  if (n1 < 0 || n2 < 0 || D < 1) return 0;
  return align_up((size_t)n1 * (size_t)n2 * 4) + (size_t)n1 * (size_t)(D + 1) * 8;
}

extern "C" int gpblur_rbf_covariance_backward(const float* x1, const float* x2, long long n1, long long n2, int D,
                                              const float* raw_lengthscale, int ard, const float* raw_outputscale,
                                              const float* g_out, float* g_x1, float* g_x2, float* g_raw_lengthscale,
                                              float* g_raw_outputscale, void* scratch, size_t scratch_bytes,
                                              void* stream) {
  if (n1 < 1 || n2 < 1 || D < 1 || D > GPBLUR_MAX_D) return GPBLUR_EINVAL;
  if (!x1 || !x2 || !raw_lengthscale || !raw_outputscale || !g_out || !g_raw_lengthscale || !g_raw_outputscale || !scratch)
    return GPBLUR_EINVAL;
  if ((reinterpret_cast<uintptr_t>(scratch) & 255) || scratch_bytes < gpblur_rbf_covariance_backward_scratch_bytes(n1, n2, D))
    return GPBLUR_EWORKSPACE;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* W = static_cast<float*>(scratch);
  double* rowstat = reinterpret_cast<double*>(static_cast<char*>(scratch) + align_up((size_t)n1 * (size_t)n2 * 4));
  const int threads = round_up(D + 1, 32);
  ProfScope ps(ST_OTHER, st);
  dim3 grid((unsigned)((n2 + 15) / 16), (unsigned)((n1 + 15) / 16));
  rbf_w_kernel<<<grid, 256, 0, st>>>(x1, x2, n1, n2, D, raw_lengthscale, ard, raw_outputscale, g_out, W);
  rbf_row_kernel<<<(unsigned)n1, threads, 0, st>>>(x1, x2, n2, D, raw_lengthscale, ard, W, g_x1, rowstat);
  if (g_x2) rbf_col_kernel<<<(unsigned)n2, threads, 0, st>>>(x1, x2, n1, n2, D, raw_lengthscale, ard, W, g_x2);
  rbf_finish_kernel<<<1, threads, 0, st>>>(rowstat, n1, D, raw_lengthscale, ard, raw_outputscale, g_raw_lengthscale,
                                           g_raw_outputscale);
  note_launch(g_x2 ? 4 : 3);
  return check_launch("rbf_cov_backward");
}
