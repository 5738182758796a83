// Micro-benchmarks of a few primitives the per-graph kernels lean on (one CTA of 512 threads).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lat lat.cu && ./lat
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void k(unsigned long long* out, int* sink) {
  __shared__ int sh[4096];
  const int t = threadIdx.x, lane = t & 31;
  for (int i = t; i < 4096; i += blockDim.x) sh[i] = i;
  __syncthreads();
  unsigned long long c0, c1;
  int acc = 0;
  // 1. match.any with 32 distinct values / 4 distinct values
  c0 = clock64();
  for (int i = 0; i < 16; ++i) acc += __popc(__match_any_sync(0xffffffffu, lane + i + acc));
  c1 = clock64();
  if (t == 0) out[0] = (c1 - c0) / 16;
  c0 = clock64();
  for (int i = 0; i < 16; ++i) acc += __popc(__match_any_sync(0xffffffffu, ((lane + acc) & 3) + i));
  c1 = clock64();
  if (t == 0) out[1] = (c1 - c0) / 16;
  // 2. __syncthreads, 512 threads
  __syncthreads();
  c0 = clock64();
  for (int i = 0; i < 16; ++i) __syncthreads();
  c1 = clock64();
  if (t == 0) out[2] = (c1 - c0) / 16;
  // 3. dependent shared-memory load chain
  int p = t & 1023;
  c0 = clock64();
  for (int i = 0; i < 16; ++i) p = sh[(p * 7 + 1) & 4095];
  c1 = clock64();
  if (t == 0) out[3] = (c1 - c0) / 16;
  acc += p;
  // 4. shared atomicAdd, distinct addresses / all lanes same address
  c0 = clock64();
  for (int i = 0; i < 16; ++i) acc += atomicAdd(&sh[(t + i * 32) & 4095], 1);
  c1 = clock64();
  if (t == 0) out[4] = (c1 - c0) / 16;
  c0 = clock64();
  for (int i = 0; i < 16; ++i) acc += atomicAdd(&sh[(t >> 5) + i], 1);
  c1 = clock64();
  if (t == 0) out[5] = (c1 - c0) / 16;
  // 5. warp shuffle dependent chain
  c0 = clock64();
  for (int i = 0; i < 16; ++i) acc += __shfl_xor_sync(0xffffffffu, acc, 1 + (i & 15));
  c1 = clock64();
  if (t == 0) out[6] = (c1 - c0) / 16;
  // 6. __reduce_add_sync
  c0 = clock64();
  for (int i = 0; i < 16; ++i) acc += __reduce_add_sync(0xffffffffu, acc);
  c1 = clock64();
  if (t == 0) out[7] = (c1 - c0) / 16;
  // 7. global load dependent chain (L2 hit after first touch)
  c0 = clock64();
  int q = t;
  for (int i = 0; i < 8; ++i) q = sink[(q * 33 + 7) & 65535];
  c1 = clock64();
  if (t == 0) out[8] = (c1 - c0) / 8;
  acc += q;
  c0 = clock64();
  for (int i = 0; i < 8; ++i) q = sink[(q * 33 + 7) & 65535];
  c1 = clock64();
  if (t == 0) out[9] = (c1 - c0) / 8;
  acc += q;
  sink[65536 + t] = acc;
}
int main() {
  unsigned long long* out; int* sink;
  cudaMalloc(&out, 16 * 8); cudaMalloc(&sink, (65536 + 1024) * 4);
  cudaMemset(sink, 0, (65536 + 1024) * 4);
  for (int it = 0; it < 2; ++it) k<<<1, 512>>>(out, sink);
  unsigned long long h[16];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const char* names[] = {"match.any 32 distinct", "match.any 4 distinct", "__syncthreads (512 thr)", "dependent LDS", "shared atomicAdd distinct",
                         "shared atomicAdd same addr per warp", "dependent SHFL+add", "__reduce_add_sync", "dependent global load (1st)", "dependent global load (2nd)"};
  for (int i = 0; i < 10; ++i) printf("%-40s %llu cycles\n", names[i], h[i]);
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
