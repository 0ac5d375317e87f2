// Development probe: per-node latency of a captured chain of short kernels with and without programmatic dependent launch.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("ERR %s line %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <bool PDL, bool EARLY>
__global__ void k(float* a, int n, int spin) {
  __shared__ float s[256];
  s[threadIdx.x] = 0.f;                      // "prologue"
  if (EARLY) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  if (PDL) asm volatile("griddepcontrol.wait;" ::: "memory");
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  float v = a[i % n];
  for (int j = 0; j < spin; ++j) v = v * 1.0001f + 0.5f;
  a[i % n] = v + s[threadIdx.x];
}

template <bool PDL, bool EARLY>
int run(const char* name, int grid, int spin, int chain) {
  float* a; CK(cudaMalloc(&a, 1 << 22)); CK(cudaMemset(a, 0, 1 << 22));
  cudaStream_t st; CK(cudaStreamCreate(&st));
  cudaGraph_t g; cudaGraphExec_t ge;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  for (int c = 0; c < chain; ++c) {
    cudaLaunchConfig_t cfg = {}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = PDL ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, k<PDL, EARLY>, a, 1 << 20, spin));
  }
  CK(cudaStreamEndCapture(st, &g));
  CK(cudaGraphInstantiate(&ge, g, 0));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int w = 0; w < 5; ++w) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e0, st));
  const int reps = 50;
  for (int r = 0; r < reps; ++r) CK(cudaGraphLaunch(ge, st));
  CK(cudaEventRecord(e1, st));
  CK(cudaStreamSynchronize(st));
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-28s grid %5d spin %5d : %.3f us per node\n", name, grid, spin, ms * 1000.f / (reps * chain));
  return 0;
}

int main() {
  for (int grid : {148, 592, 2368}) for (int spin : {0, 2000}) {
    if (run<false, false>("plain", grid, spin, 40)) return 1;
    if (run<true, false>("pdl (implicit trigger)", grid, spin, 40)) return 1;
    if (run<true, true>("pdl + early trigger", grid, spin, 40)) return 1;
  }
  return 0;
}
