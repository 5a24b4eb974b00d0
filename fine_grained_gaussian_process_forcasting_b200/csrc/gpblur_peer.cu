// One-shot all-reduce of the flat gradient bucket over NVLink peer memory (one node, one process per GPU).
//
// The data-parallel step of this path has ONE exchange: the mean over ranks of M*D + 2M + 2D + 3 floats (17 k floats
// at the reference shape).  NCCL moves that in ~30 us inside a captured step - all of it latency.  Here every rank
// owns a small communication buffer that its peers map through CUDA IPC:
//
//   [ flags: one uint64 per rank | counters | staging 0 | staging 1 ]
//
// and one kernel per step does: copy the bucket into staging[step & 1] -> release-store (step + 1) into MY slot of
// every peer's flag array -> wait until all peers have stored theirs in mine -> read every rank's staging buffer
// (peer loads over NVLink) and sum in RANK ORDER -> scale -> write the bucket.  Every rank adds the same numbers in
// the same order: the result is bit-identical on all ranks and independent of timing.
//
// Why two staging buffers are enough: a rank cannot finish step s before every peer has signalled step s, and a peer
// cannot signal step s + 1 before it has finished step s, so while I read staging[s & 1] nobody can be further than
// step s + 1 - which writes the OTHER buffer.
//
// Several CTAs share the work; they are sequenced with two monotonic device counters (arrivals of the copy phase /
// of the reduce phase), and the last CTA of the reduce phase advances the device-resident step word, so the kernel
// can be replayed from a CUDA graph without host involvement.  Waits give up (trap) after ~4 s instead of hanging the
// GPU when a peer never arrives.
#include <cstring>

#include "gpblur_common.cuh"

namespace gpblur {

namespace {

constexpr int kPeerMaxWorld = 16;
constexpr int kPeerThreads = 512;
constexpr size_t kPeerHeader = 1024;   // flags [0, 128) | cntA 128 | cntB 136 | step 144

struct PeerArgs {
  float* bucket;
  long long n;
  int world, rank;
  float scale;
  unsigned char* comm[kPeerMaxWorld];
  size_t staging_bytes;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ float4 ld_volatile4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}
__device__ __forceinline__ float ld_volatile1(const float* p) {
  float v;
  asm volatile("ld.volatile.global.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_kernel(PeerArgs a) {
  unsigned char* mine = a.comm[a.rank];
  unsigned long long* flags = reinterpret_cast<unsigned long long*>(mine);
  unsigned long long* cntA = reinterpret_cast<unsigned long long*>(mine + 128);
  unsigned long long* cntB = reinterpret_cast<unsigned long long*>(mine + 136);
  unsigned long long* stepw = reinterpret_cast<unsigned long long*>(mine + 144);
  const int tid = threadIdx.x, nct = gridDim.x;
  __shared__ unsigned long long step_s;
  if (tid == 0) step_s = *reinterpret_cast<volatile unsigned long long*>(stepw);
  __syncthreads();
  const unsigned long long step = step_s;
  const size_t soff = kPeerHeader + (size_t)(step & 1) * a.staging_bytes;
  float* my_stage = reinterpret_cast<float*>(mine + soff);
  // slice of this CTA (multiples of 4 floats)
  const long long n4 = (a.n + 3) / 4;
  const long long per = (n4 + nct - 1) / nct;
  const long long q0 = (long long)blockIdx.x * per, q1 = q0 + per < n4 ? q0 + per : n4;

  // ---- phase A: bucket -> my staging buffer ----
  for (long long q = q0 + tid; q < q1; q += kPeerThreads) {
    const long long e = 4 * q;
    if (e + 3 < a.n) {
      *reinterpret_cast<float4*>(my_stage + e) = *reinterpret_cast<const float4*>(a.bucket + e);
    } else {
      for (long long k = e; k < a.n; ++k) my_stage[k] = a.bucket[k];
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    const unsigned long long old = atomicAdd(cntA, 1ull);
    if (old == (step + 1) * (unsigned long long)nct - 1) {
      // every CTA of this rank has staged its slice: tell the peers (and myself)
      // ONE system-scope fence, then plain volatile stores (a release per peer made the fences queue up: 8 ranks paid
      // ~50 us for their 8 signals)
      __threadfence_system();
      for (int p = 0; p < a.world; ++p)
        *reinterpret_cast<volatile unsigned long long*>(reinterpret_cast<unsigned long long*>(a.comm[p]) + a.rank) = step + 1;
    }
  }
  // ---- wait for every rank's signal (thread q watches rank q) ----
  if (tid < a.world) {
    const long long t0 = clock64();
    while (ld_acquire_sys(flags + tid) < step + 1) {
      if (clock64() - t0 > (1ll << 33)) __trap();          // ~4 s: a peer never arrived
    }
  }
  __syncthreads();

  // ---- phase B: sum the staging buffers in rank order ----
  for (long long q = q0 + tid; q < q1; q += kPeerThreads) {
    const long long e = 4 * q;
    if (e + 3 < a.n) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int p = 0; p < a.world; ++p) {
        const float4 v = ld_volatile4(reinterpret_cast<const float*>(a.comm[p] + soff) + e);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      acc.x *= a.scale; acc.y *= a.scale; acc.z *= a.scale; acc.w *= a.scale;
      *reinterpret_cast<float4*>(a.bucket + e) = acc;
    } else {
      for (long long k = e; k < a.n; ++k) {
        float acc = 0.f;
        for (int p = 0; p < a.world; ++p) acc += ld_volatile1(reinterpret_cast<const float*>(a.comm[p] + soff) + k);
        a.bucket[k] = acc * a.scale;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned long long old = atomicAdd(cntB, 1ull);
    if (old == (step + 1) * (unsigned long long)nct - 1) {
      __threadfence();
      *reinterpret_cast<volatile unsigned long long*>(stepw) = step + 1;
    }
  }
}

}  // namespace

}  // namespace gpblur

using namespace gpblur;

extern "C" {

size_t gpblur_peer_comm_bytes(long long n) {
  if (n < 0) return 0;
  const size_t staging = align_up((size_t)((n + 3) / 4 * 4) * sizeof(float), 256);
  return kPeerHeader + 2 * staging;
}

int gpblur_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  if (!ptr || !handle64 || bytes == 0) return GPBLUR_EINVAL;
  void* p = nullptr;
  if (cudaMalloc(&p, bytes) != cudaSuccess) return GPBLUR_ELAUNCH;
  if (cudaMemset(p, 0, bytes) != cudaSuccess) return GPBLUR_ELAUNCH;
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); return GPBLUR_ELAUNCH; }
  static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(handle64, &h, 64);
  if (cudaDeviceSynchronize() != cudaSuccess) return GPBLUR_ELAUNCH;
  *ptr = p;
  return GPBLUR_OK;
}

int gpblur_peer_open(const unsigned char* handle64, void** ptr) {
  if (!ptr || !handle64) return GPBLUR_EINVAL;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  void* p = nullptr;
  if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
    cudaGetLastError();
    return GPBLUR_ELAUNCH;
  }
  *ptr = p;
  return GPBLUR_OK;
}

int gpblur_peer_close(void* ptr) { return cudaIpcCloseMemHandle(ptr) == cudaSuccess ? GPBLUR_OK : GPBLUR_ELAUNCH; }
int gpblur_peer_free(void* ptr) { return cudaFree(ptr) == cudaSuccess ? GPBLUR_OK : GPBLUR_ELAUNCH; }

int gpblur_peer_allreduce(float* bucket, long long n, int world, int rank, void* const* comm, float scale,
                          void* stream) {
  if (!bucket || n < 0 || world < 1 || world > kPeerMaxWorld || rank < 0 || rank >= world || !comm) return GPBLUR_EINVAL;
  if (reinterpret_cast<uintptr_t>(bucket) & 15) return GPBLUR_EINVAL;
  if (n == 0) return GPBLUR_OK;
  PeerArgs a{};
  a.bucket = bucket; a.n = n; a.world = world; a.rank = rank; a.scale = scale;
  for (int p = 0; p < world; ++p) {
    if (!comm[p]) return GPBLUR_EINVAL;
    a.comm[p] = reinterpret_cast<unsigned char*>(comm[p]);
  }
  a.staging_bytes = align_up((size_t)((n + 3) / 4 * 4) * sizeof(float), 256);
  // the exchange is latency, not bandwidth: one float4 per thread and peer where possible (one NVLink round trip),
  // at most 16 CTAs (all co-resident: they synchronise through device counters)
  int grid = (int)((n + 4 * kPeerThreads - 1) / (4 * kPeerThreads));
  if (grid > 16) grid = 16;
  peer_allreduce_kernel<<<grid, kPeerThreads, 0, (cudaStream_t)stream>>>(a);
  note_launch();
  return check_launch("peer_allreduce");
}

}  // extern "C"
