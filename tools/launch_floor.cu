// Micro-benchmark (run under gpurun): the floor of a dependent kernel chain on this GPU.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/launch_floor tools/launch_floor.cu && /tmp/launch_floor
// Reports microseconds per kernel node for chains replayed as one CUDA graph:
//   empty kernels, kernels with one dependent L2 round trip, with / without programmatic dependent
//   launch edges, at a few grid sizes.  The segmentation's 40 relabelling passes are such a chain.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

__global__ void k_empty(int* p, int pdl) {
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p == nullptr) printf("never\n");
}
// one load that depends on a value the previous kernel wrote, one store
__global__ void k_chain(int* p, int n, int pdl) {
  if (pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = p[(i * 7 + 13) % n] + 1;
}

template <typename F>
static float time_graph(cudaStream_t st, int nodes, int reps, F enqueue) {
  cudaGraph_t g;
  cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  for (int i = 0; i < nodes; i++) enqueue();
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaGraphUpload(ge, st);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 3; i++) cudaGraphLaunch(ge, st);
  cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  for (int i = 0; i < reps; i++) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st);
  cudaStreamSynchronize(st);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return ms * 1e3f / (reps * nodes);
}

static void launch(cudaStream_t st, bool pdl, void (*k)(int*, int, int), int grid, int block, int* p, int n) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k, p, n, pdl ? 1 : 0);
}
static void launch_e(cudaStream_t st, bool pdl, int grid, int block, int* p) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k_empty, p, pdl ? 1 : 0);
}

int main() {
  cudaStream_t st;
  cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  int* p;
  const int n = 1 << 20;
  cudaMalloc(&p, n * sizeof(int));
  cudaMemset(p, 0, n * sizeof(int));
  printf("{\"what\": \"us per kernel node of a dependent chain replayed as a CUDA graph (100 nodes x 50 replays)\"");
  for (int pdl = 0; pdl < 2; pdl++) {
    for (int grid : {1, 148, 330, 1184}) {
      const float e = time_graph(st, 100, 50, [&] { launch_e(st, pdl, grid, 256, p); });
      const float c = time_graph(st, 100, 50, [&] { launch(st, pdl, k_chain, grid, 256, p, grid * 256); });
      printf(", \"%s_grid%d\": {\"empty\": %.3f, \"one_l2_round_trip\": %.3f}", pdl ? "pdl" : "plain", grid, e, c);
    }
  }
  printf("}\n");
  return 0;
}
