// Does a programmatic dependent launch overlap a long primary (128 CTAs, 200 KB smem) with a secondary (64 CTAs)?
// eager and captured into a CUDA graph.  nvcc -arch=sm_100a pdl_test.cu -o pdl_test
#include <cstdio>
#include <cuda_runtime.h>
__global__ void primary(long long ns, int trigger) {
  extern __shared__ float sm[];
  if (trigger) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  long long t1 = t0;
  while (t1 - t0 < ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  sm[threadIdx.x] = (float)t1;
}
__global__ void secondary(long long ns, int wait) {
  extern __shared__ float sm[];
  long long t0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  long long t1 = t0;
  while (t1 - t0 < ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
  sm[threadIdx.x] = (float)t1;
  if (wait) asm volatile("griddepcontrol.wait;" ::: "memory");
}
static int PG = 128, PS = 200 * 1024, SS = 88 * 1024, ST = 512;
static void launch_pair(cudaStream_t st, int pdl, int trigger) {
  primary<<<PG, 512, PS, st>>>(20000, trigger);
  if (pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(64); cfg.blockDim = dim3(ST); cfg.dynamicSmemBytes = SS; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, secondary, (long long)8000, 1);
  } else {
    secondary<<<64, ST, SS, st>>>(8000, 0);
  }
}
int main() {
  cudaFuncSetAttribute(primary, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(secondary, cudaFuncAttributeMaxDynamicSharedMemorySize, 88 * 1024);
  cudaStream_t st; cudaStreamCreate(&st);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int var = 0; var < 3; ++var) {
  if (var == 1) { SS = 0; ST = 128; }
  if (var == 2) { PG = 64; PS = 0; }
  printf("variant %d: primary %d CTAs smem %d, secondary 64 CTAs x %d threads smem %d\n", var, PG, PS, ST, SS);
  for (int mode = 0; mode < 3; mode += 2) {          // 0 plain, 1 pdl without trigger, 2 pdl with trigger
    const int pdl = mode > 0, trig = mode == 2;
    for (int graph = 0; graph < 2; ++graph) {
      cudaGraphExec_t ge = nullptr;
      if (graph) {
        cudaGraph_t g;
        cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
        for (int i = 0; i < 10; ++i) launch_pair(st, pdl, trig);
        cudaStreamEndCapture(st, &g);
        cudaError_t e = cudaGraphInstantiate(&ge, g, 0);
        if (e != cudaSuccess) { printf("instantiate: %s\n", cudaGetErrorString(e)); return 1; }
      }
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a, st);
        if (graph) for (int i = 0; i < 10; ++i) cudaGraphLaunch(ge, st);
        else for (int i = 0; i < 100; ++i) launch_pair(st, pdl, trig);
        cudaEventRecord(b, st);
        cudaStreamSynchronize(st);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (rep) printf("mode %d (%s) graph %d: %.2f us per pair (primary 20 us, secondary 8 us)   %s\n", mode,
                        mode == 0 ? "plain" : mode == 1 ? "pdl, no trigger" : "pdl + trigger", graph, 1e3 * ms / 100,
                        cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  }
  return 0;
}
