// Node-to-node latency of small dependent kernels replayed from a CUDA graph: plain stream order vs programmatic
// dependent launch (griddepcontrol.wait first in every kernel).  nvcc -arch=sm_100a -O3 pdl_gap.cu -o pdl_gap
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_plain(float* p, int work) {
  float v = p[threadIdx.x];
  for (int i = 0; i < work; ++i) v = v * 1.0001f + 0.5f;
  p[threadIdx.x] = v;
}
__global__ void k_pdl(float* p, int work) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  float v = p[threadIdx.x];
  for (int i = 0; i < work; ++i) v = v * 1.0001f + 0.5f;
  p[threadIdx.x] = v;
}

static float run(bool pdl, int nodes, int ctas, int work, float* d, cudaStream_t st) {
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < nodes; ++i) {
    if (!pdl) {
      k_plain<<<ctas, 256, 0, st>>>(d, work);
    } else {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      cudaLaunchKernelEx(&cfg, k_pdl, d, work);
    }
  }
  cudaError_t e = cudaStreamEndCapture(st, &g);
  if (e != cudaSuccess) { printf("capture failed: %s\n", cudaGetErrorString(e)); return -1.f; }
  e = cudaGraphInstantiate(&ge, g, 0);
  if (e != cudaSuccess) { printf("instantiate failed: %s\n", cudaGetErrorString(e)); return -1.f; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 5; ++i) cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  const int reps = 50;
  for (int i = 0; i < reps; ++i) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1e3f / (reps * nodes);
}

int main() {
  float* d;
  cudaMalloc(&d, 1 << 20);
  cudaMemset(d, 0, 1 << 20);
  cudaStream_t st;
  cudaStreamCreate(&st);
  for (int ctas : {1, 148, 592})
    for (int work : {0, 2000, 20000}) {
      const float a = run(false, 12, ctas, work, d, st), b = run(true, 12, ctas, work, d, st);
      printf("ctas %4d work %6d: plain %.2f us / node, pdl %.2f us / node\n", ctas, work, a, b);
    }
  printf("last error: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
