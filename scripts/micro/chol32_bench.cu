// Timing variants of the 32 x 32 warp Cholesky (one row per lane) on one warp.  GPU box only.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o chol32_bench chol32_bench.cu
#include <cmath>
#include <cstdio>
#include <cuda_runtime.h>
constexpr int TB = 32;
__device__ __forceinline__ double rsq(double d) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  const double e = fma(-d * y, y, 1.0);
  return fma(y * e, fma(0.375, e, 0.5), y);
}
// V = 0: column through shared memory (current code, no fence); 1: chain only (no trailing update of columns > c + 1);
// 2: multipliers by shuffle instead of shared memory; 3: shared memory, trailing update split in "urgent" (c+1, c+2) and rest
template <int V>
__device__ __forceinline__ void chol(double (&a)[TB], int lane, double* Lt) {
  double d = __shfl_sync(0xffffffffu, a[0], 0);
  double rs = rsq(d);
#pragma unroll
  for (int c = 0; c < TB; ++c) {
    const double l = a[c] * rs;
    if (c + 1 < TB) {
      const double piv = fma(-l, l, a[c + 1]);
      d = __shfl_sync(0xffffffffu, piv, c + 1);
      rs = rsq(d);
    }
    double* col = Lt + c * TB;
    if (V == 2) {
      a[c] = l;
#pragma unroll
      for (int c2 = c + 1; c2 < TB; ++c2) {
        const double m = __shfl_sync(0xffffffffu, l, c2);
        a[c2] = fma(-l, m, a[c2]);
      }
    } else {
      col[lane] = (lane >= c) ? l : 0.0;
      __syncwarp();
      if (V == 1) {
        if (c + 1 < TB) a[c + 1] = fma(-l, col[c + 1], a[c + 1]);
      } else {
        if ((c + 1) & 1) {
          if (c + 1 < TB) a[c + 1] = fma(-l, col[c + 1], a[c + 1]);
#pragma unroll
          for (int c2 = c + 2; c2 + 1 < TB; c2 += 2) {
            const double2 m2 = *reinterpret_cast<const double2*>(col + c2);
            a[c2] = fma(-l, m2.x, a[c2]);
            a[c2 + 1] = fma(-l, m2.y, a[c2 + 1]);
          }
        } else {
#pragma unroll
          for (int c2 = c + 1; c2 + 1 < TB; c2 += 2) {
            const double2 m2 = *reinterpret_cast<const double2*>(col + c2);
            a[c2] = fma(-l, m2.x, a[c2]);
            a[c2 + 1] = fma(-l, m2.y, a[c2 + 1]);
          }
        }
      }
    }
  }
}
template <int V>
__global__ void k(const double* A, double* L, long long* cyc) {
  __shared__ __align__(16) double Lt[TB * TB];
  const int lane = threadIdx.x;
  double a[TB];
#pragma unroll
  for (int c = 0; c < TB; ++c) a[c] = A[lane * TB + c];
  __syncwarp();
  const long long t0 = clock64();
  chol<V>(a, lane, Lt);
  double s = 0;
#pragma unroll
  for (int c = 0; c < TB; ++c) s += a[c];
  const long long t1 = clock64() + (long long)(s == 1.2345);
  if (lane == 0) cyc[V] = t1 - t0;
  __syncwarp();
  if (V != 2)
    for (int c = 0; c < TB; ++c) L[lane * TB + c] = Lt[c * TB + lane];
  else
    for (int c = 0; c < TB; ++c) L[lane * TB + c] = lane >= c ? a[c] : 0.0;
}

// ---- two warps: warp 0 factor (publishes columns + flag), warp 1 inverse.  F = 0: plain volatile flag, 1: fence.acq_rel.cta,
// 2: __threadfence_block (fence.sc.cta), 3: st.release.cta / ld.acquire.cta
template <int F>
__device__ __forceinline__ void publish(volatile int* ready, int v) {
  if (F == 1) asm volatile("fence.acq_rel.cta;" ::: "memory");
  if (F == 2) __threadfence_block();
  if (F == 3) asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared((const void*)ready)), "r"(v) : "memory");
  else *ready = v;
}
template <int F>
__device__ __forceinline__ void wait_for(volatile int* ready, int c) {
  if (F == 3) {
    int v;
    do { asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((unsigned)__cvta_generic_to_shared((const void*)ready)) : "memory"); } while (v <= c);
  } else {
    while (*ready <= c) { }
    if (F == 1) asm volatile("fence.acq_rel.cta;" ::: "memory");
    if (F == 2) __threadfence_block();
    if (F == 0) asm volatile("" ::: "memory");
  }
}
template <int F>
__global__ void k2(const double* A, double* L, double* X, long long* cyc) {
  __shared__ __align__(16) double Lt[TB * TB];
  __shared__ double rsv[TB];
  __shared__ volatile int ready;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) ready = 0;
  double a[TB];
#pragma unroll
  for (int c = 0; c < TB; ++c) a[c] = A[lane * TB + c];
  __syncthreads();
  const long long t0 = clock64();
  if (warp == 0) {
    double d = __shfl_sync(0xffffffffu, a[0], 0);
    double rs = rsq(d);
#pragma unroll
    for (int c = 0; c < TB; ++c) {
      const double l = a[c] * rs;
      const double rs_c = rs;
      if (c + 1 < TB) {
        const double piv = fma(-l, l, a[c + 1]);
        d = __shfl_sync(0xffffffffu, piv, c + 1);
        rs = rsq(d);
      }
      double* col = Lt + c * TB;
      col[lane] = (lane >= c) ? l : 0.0;
      if (lane == 0) rsv[c] = rs_c;
      __syncwarp();
      if (lane == 0) publish<F>(&ready, c + 1);
      if ((c + 1) & 1) {
        if (c + 1 < TB) a[c + 1] = fma(-l, col[c + 1], a[c + 1]);
#pragma unroll
        for (int c2 = c + 2; c2 + 1 < TB; c2 += 2) {
          const double2 m2 = *reinterpret_cast<const double2*>(col + c2);
          a[c2] = fma(-l, m2.x, a[c2]);
          a[c2 + 1] = fma(-l, m2.y, a[c2 + 1]);
        }
      } else {
#pragma unroll
        for (int c2 = c + 1; c2 + 1 < TB; c2 += 2) {
          const double2 m2 = *reinterpret_cast<const double2*>(col + c2);
          a[c2] = fma(-l, m2.x, a[c2]);
          a[c2 + 1] = fma(-l, m2.y, a[c2 + 1]);
        }
      }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < TB; ++c) s += a[c];
    const long long t1 = clock64() + (long long)(s == 1.2345);
    if (lane == 0) cyc[0] = t1 - t0;
  } else {
    double x[TB];
#pragma unroll
    for (int r = 0; r < TB; ++r) x[r] = (r == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int c = 0; c < TB; ++c) {
      wait_for<F>(&ready, c);
      const double* col = Lt + c * TB;
      const double xc = (c >= lane) ? x[c] * rsv[c] : 0.0;
      x[c] = xc;
      if ((c + 1) & 1) {
        if (c + 1 < TB) x[c + 1] = fma(-col[c + 1], xc, x[c + 1]);
#pragma unroll
        for (int c2 = c + 2; c2 + 1 < TB; c2 += 2) {
          const double2 m2 = *reinterpret_cast<const double2*>(col + c2);
          x[c2] = fma(-m2.x, xc, x[c2]);
          x[c2 + 1] = fma(-m2.y, xc, x[c2 + 1]);
        }
      } else {
#pragma unroll
        for (int c2 = c + 1; c2 + 1 < TB; c2 += 2) {
          const double2 m2 = *reinterpret_cast<const double2*>(col + c2);
          x[c2] = fma(-m2.x, xc, x[c2]);
          x[c2 + 1] = fma(-m2.y, xc, x[c2 + 1]);
        }
      }
    }
    double s = 0;
#pragma unroll
    for (int c = 0; c < TB; ++c) s += x[c];
    const long long t1 = clock64() + (long long)(s == 1.2345);
    if (lane == 0) cyc[1] = t1 - t0;
#pragma unroll
    for (int r = 0; r < TB; ++r) X[r * TB + lane] = x[r];
  }
  __syncthreads();
  if (warp == 0)
    for (int c = 0; c < TB; ++c) L[lane * TB + c] = Lt[c * TB + lane];
}

// ---- one warp, factor + inverse fused in the same column loop (the round-1 kernel); B: with bad-pivot tracking
template <int B>
__global__ void k3(const double* A, double* L, double* X, long long* cyc, int* info) {
  __shared__ __align__(16) double lcol[2 * TB];
  const int lane = threadIdx.x;
  double a[TB], x[TB];
#pragma unroll
  for (int c = 0; c < TB; ++c) a[c] = A[lane * TB + c];
  __syncwarp();
  unsigned long long n0;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(n0));
  const long long t0 = clock64();
  int bad = 0;
#pragma unroll
  for (int r = 0; r < TB; ++r) x[r] = (r == lane) ? 1.0 : 0.0;
  double d = __shfl_sync(0xffffffffu, a[0], 0);
  if (B && !(d > 0.0)) bad = 1;
  double rs = rsq(d);
#pragma unroll
  for (int c = 0; c < TB; ++c) {
    const double l = a[c] * rs;
    const double xc = (c >= lane) ? x[c] * rs : 0.0;
    if (c + 1 < TB) {
      const double piv = fma(-l, l, a[c + 1]);
      d = __shfl_sync(0xffffffffu, piv, c + 1);
      if (B && !(d > 0.0) && bad == 0) bad = c + 2;
      rs = rsq(d);
    }
    a[c] = (lane >= c) ? l : 0.0;
    x[c] = xc;
    double* buf = lcol + (c & 1) * TB;
    buf[lane] = l;
    __syncwarp();
    if ((c + 1) & 1) {
      if (c + 1 < TB) {
        const double m = buf[c + 1];
        a[c + 1] = fma(-l, m, a[c + 1]);
        x[c + 1] = fma(-m, xc, x[c + 1]);
      }
#pragma unroll
      for (int c2 = c + 2; c2 + 1 < TB; c2 += 2) {
        const double2 m2 = *reinterpret_cast<const double2*>(buf + c2);
        a[c2] = fma(-l, m2.x, a[c2]);
        a[c2 + 1] = fma(-l, m2.y, a[c2 + 1]);
        x[c2] = fma(-m2.x, xc, x[c2]);
        x[c2 + 1] = fma(-m2.y, xc, x[c2 + 1]);
      }
    } else {
#pragma unroll
      for (int c2 = c + 1; c2 + 1 < TB; c2 += 2) {
        const double2 m2 = *reinterpret_cast<const double2*>(buf + c2);
        a[c2] = fma(-l, m2.x, a[c2]);
        a[c2 + 1] = fma(-l, m2.y, a[c2 + 1]);
        x[c2] = fma(-m2.x, xc, x[c2]);
        x[c2 + 1] = fma(-m2.y, xc, x[c2 + 1]);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < TB; ++c) s += a[c] + x[c];
  const long long t1 = clock64() + (long long)(s == 1.2345);
  unsigned long long n1;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(n1));
  if (lane == 0) { cyc[0] = t1 - t0; cyc[1] = t1 - t0; cyc[2] = (long long)(n1 - n0); if (bad) info[0] = bad; }
#pragma unroll
  for (int c = 0; c < TB; ++c) { L[lane * TB + c] = a[c]; X[c * TB + lane] = x[c]; }
}
int main() {
  double hA[TB * TB], hL[TB * TB], ref[TB * TB];
  for (int i = 0; i < TB; ++i)
    for (int j = 0; j < TB; ++j) hA[i * TB + j] = exp(-0.05 * (i - j) * (i - j)) + (i == j ? 1e-2 : 0.0);
  for (int i = 0; i < TB * TB; ++i) ref[i] = 0;
  for (int j = 0; j < TB; ++j) {
    double s = hA[j * TB + j];
    for (int k2 = 0; k2 < j; ++k2) s -= ref[j * TB + k2] * ref[j * TB + k2];
    ref[j * TB + j] = sqrt(s);
    for (int i = j + 1; i < TB; ++i) {
      double t = hA[i * TB + j];
      for (int k2 = 0; k2 < j; ++k2) t -= ref[i * TB + k2] * ref[j * TB + k2];
      ref[i * TB + j] = t / ref[j * TB + j];
    }
  }
  double *dA, *dL; long long* cyc;
  cudaMalloc(&dA, sizeof(hA)); cudaMalloc(&dL, sizeof(hA)); cudaMalloc(&cyc, 64);
  cudaMemcpy(dA, hA, sizeof(hA), cudaMemcpyHostToDevice);
  long long h[8];
  for (int v = 0; v < 3; ++v) {
    for (int rep = 0; rep < 3; ++rep) {
      if (v == 0) k<0><<<1, 32>>>(dA, dL, cyc);
      if (v == 1) k<1><<<1, 32>>>(dA, dL, cyc);
      if (v == 2) k<2><<<1, 32>>>(dA, dL, cyc);
    }
    cudaMemcpy(hL, dL, sizeof(hA), cudaMemcpyDeviceToHost);
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    double err = 0;
    for (int i = 0; i < TB * TB; ++i) err = fmax(err, fabs(hL[i] - ref[i]));
    printf("variant %d: %lld cycles (%.1f per column), max err vs host %.2e  [%s]\n", v, h[v], h[v] / 32.0, err, cudaGetErrorString(cudaGetLastError()));
  }
  double* dX; cudaMalloc(&dX, sizeof(hA));
  double hX[TB * TB];
  int* dinfo; cudaMalloc(&dinfo, 4);
  for (int f = 0; f < 6; ++f) {
    for (int rep = 0; rep < 3; ++rep) {
      if (f == 4) k3<0><<<1, 32>>>(dA, dL, dX, cyc, dinfo);
      if (f == 5) k3<1><<<1, 32>>>(dA, dL, dX, cyc, dinfo);
      if (f == 0) k2<0><<<1, 64>>>(dA, dL, dX, cyc);
      if (f == 1) k2<1><<<1, 64>>>(dA, dL, dX, cyc);
      if (f == 2) k2<2><<<1, 64>>>(dA, dL, dX, cyc);
      if (f == 3) k2<3><<<1, 64>>>(dA, dL, dX, cyc);
    }
    cudaMemcpy(hL, dL, sizeof(hA), cudaMemcpyDeviceToHost);
    cudaMemcpy(hX, dX, sizeof(hA), cudaMemcpyDeviceToHost);
    cudaMemcpy(h, cyc, 64, cudaMemcpyDeviceToHost);
    double err = 0, errx = 0;
    for (int i = 0; i < TB * TB; ++i) err = fmax(err, fabs(hL[i] - ref[i]));
    for (int i = 0; i < TB; ++i)
      for (int j = 0; j < TB; ++j) {   // L X = I
        double t = 0;
        for (int k3 = 0; k3 < TB; ++k3) t += ref[i * TB + k3] * hX[k3 * TB + j];
        errx = fmax(errx, fabs(t - (i == j ? 1.0 : 0.0)));
      }
    if (f >= 4) printf("  (globaltimer: %lld ns) ", h[2]);
    printf("two warps, flag mode %d: factor %lld cycles, inverse done at %lld, err L %.2e, |L X - I| %.2e [%s]\n", f, h[0], h[1], err, errx,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
