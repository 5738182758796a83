// Cluster max-pool and graph read-out.
//
// max-pool  : torch_scatter.scatter_max as used by community_pooling
//             (deeprank_gnn/community_pooling.py:201) and PyG max_pool_x
//             (ginet.py:114,129, sGAT.py:130, foutnet.py:117).  Members of a cluster come from the
//             cluster -> members CSR of the structure pass (ascending node id), so the CPU tie
//             rule "first occurrence wins" (update only on strict >) falls out of the scan order
//             and no atomics are needed.  A NaN never wins (NaN > x is false); a segment in which
//             nothing beats the initial value yields 0 with argmax = -1.
// read-out  : torch_scatter.scatter_mean(x, batch) (ginet.py:133-134, sGAT.py:133, foutnet.py:120).
#include <float.h>

#include "common.cuh"

namespace drgnn {

// one thread per (cluster, 4 channels)
__global__ void __launch_bounds__(256) maxpool_fwd_vec_kernel(const float* __restrict__ x, int ldx,
                                                              const int32_t* __restrict__ cmptr,
                                                              const int32_t* __restrict__ cmem, int n_clusters,
                                                              const int32_t* n_dev, int C4, float* __restrict__ y, int ldy,
                                                              int32_t* __restrict__ argmax) {
  const int n = n_dev ? min(*n_dev, n_clusters) : n_clusters;
  const int64_t total = (int64_t)n * C4;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx / C4), q = (int)(idx % C4);
    const int s = __ldg(cmptr + k), e = __ldg(cmptr + k + 1);
    float4 best = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    int4 arg = make_int4(-1, -1, -1, -1);
    for (int p = s; p < e; ++p) {
      const int i = __ldg(cmem + p);
      const float4 v = *reinterpret_cast<const float4*>(x + (int64_t)i * ldx + q * 4);
      if (v.x > best.x) { best.x = v.x; arg.x = i; }
      if (v.y > best.y) { best.y = v.y; arg.y = i; }
      if (v.z > best.z) { best.z = v.z; arg.z = i; }
      if (v.w > best.w) { best.w = v.w; arg.w = i; }
    }
    if (arg.x < 0) best.x = 0.f;
    if (arg.y < 0) best.y = 0.f;
    if (arg.z < 0) best.z = 0.f;
    if (arg.w < 0) best.w = 0.f;
    *reinterpret_cast<float4*>(y + (int64_t)k * ldy + q * 4) = best;
    *reinterpret_cast<int4*>(argmax + ((int64_t)k * C4 + q) * 4) = arg;
  }
}

// scalar fallback: one thread per (cluster, channel)
__global__ void __launch_bounds__(256) maxpool_fwd_kernel(const float* __restrict__ x, int ldx,
                                                          const int32_t* __restrict__ cmptr,
                                                          const int32_t* __restrict__ cmem, int n_clusters,
                                                          const int32_t* n_dev, int C, float* __restrict__ y, int ldy,
                                                          int32_t* __restrict__ argmax) {
  const int n = n_dev ? min(*n_dev, n_clusters) : n_clusters;
  const int64_t total = (int64_t)n * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx / C), c = (int)(idx % C);
    const int s = cmptr[k], e = cmptr[k + 1];
    float best = -FLT_MAX;
    int arg = -1;
    for (int p = s; p < e; ++p) {
      const int i = cmem[p];
      const float v = x[(int64_t)i * ldx + c];
      if (v > best) { best = v; arg = i; }
    }
    y[(int64_t)k * ldy + c] = arg < 0 ? 0.f : best;
    argmax[(int64_t)k * C + c] = arg;
  }
}

// dx[i,c] = (argmax[cl[i],c] == i) ? g[cl[i],c] : 0, optionally gated by relu_out[i,c] > 0
__global__ void __launch_bounds__(256) maxpool_bwd_kernel(const float* __restrict__ g, int ldg,
                                                          const int32_t* __restrict__ argmax,
                                                          const int32_t* __restrict__ cl,
                                                          const float* __restrict__ relu_out, int ld_relu, int n_nodes,
                                                          const int32_t* n_dev, int C, float* __restrict__ dx, int lddx) {
  const int n = n_dev ? min(*n_dev, n_nodes) : n_nodes;
  const int64_t total = (int64_t)n * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx / C), c = (int)(idx % C);
    const int k = __ldg(cl + i);
    float v = 0.f;
    if (__ldg(argmax + (int64_t)k * C + c) == i) v = __ldg(g + (int64_t)k * ldg + c);
    if (relu_out && !(relu_out[(int64_t)i * ld_relu + c] > 0.f)) v = 0.f;
    dx[(int64_t)i * lddx + c] = v;
  }
}

// one CTA per segment, threads over channels, rows summed in ascending order (the CPU
// scatter_add order)
__global__ void __launch_bounds__(256) segment_mean_fwd_kernel(const float* __restrict__ x, int ldx,
                                                               const int32_t* __restrict__ seg_ptr, int C,
                                                               float* __restrict__ r, int ldr) {
  const int b = blockIdx.x;
  const int s = seg_ptr[b], e = seg_ptr[b + 1];
  const float inv = 1.f / (float)max(e - s, 1);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    float acc = 0.f;
    for (int k = s; k < e; ++k) acc += x[(int64_t)k * ldx + c];
    r[(int64_t)b * ldr + c] = acc * inv;
  }
}

__global__ void __launch_bounds__(256) segment_mean_bwd_kernel(const float* __restrict__ g, int ldg,
                                                               const int32_t* __restrict__ seg_ptr, int C,
                                                               float* __restrict__ dx, int lddx) {
  const int b = blockIdx.x;
  const int s = seg_ptr[b], e = seg_ptr[b + 1];
  const float inv = 1.f / (float)max(e - s, 1);
  const int total = (e - s) * C;
  for (int idx = threadIdx.x; idx < total; idx += blockDim.x) {
    const int k = s + idx / C, c = idx % C;
    dx[(int64_t)k * lddx + c] = g[(int64_t)b * ldg + c] * inv;
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace drgnn

using namespace drgnn;

extern "C" int drgnn_maxpool_fwd(const float* x, int32_t ldx, const int32_t* cmptr, const int32_t* cmem,
                                 int32_t n_clusters, const int32_t* n_clusters_dev, int32_t C, float* y, int32_t ldy,
                                 int32_t* argmax, void* stream) {
  DRGNN_REQUIRE(x && cmptr && cmem && y && argmax, "maxpool_fwd: NULL pointer");
  DRGNN_REQUIRE(n_clusters >= 0 && C > 0 && ldx >= C && ldy >= C, "maxpool_fwd: bad sizes");
  if (n_clusters == 0) return DRGNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = device_info().sms;
  const bool vec = (C % 4 == 0) && (ldx % 4 == 0) && (ldy % 4 == 0) && aligned16(x) && aligned16(y) && aligned16(argmax);
  if (vec) {
    const int64_t total = (int64_t)n_clusters * (C / 4);
    const int blocks = (int)min_i64((total + 255) / 256, (int64_t)sms * 16);
    maxpool_fwd_vec_kernel<<<blocks, 256, 0, st>>>(x, ldx, cmptr, cmem, n_clusters, n_clusters_dev, C / 4, y, ldy, argmax);
  } else {
    const int64_t total = (int64_t)n_clusters * C;
    const int blocks = (int)min_i64((total + 255) / 256, (int64_t)sms * 16);
    maxpool_fwd_kernel<<<blocks, 256, 0, st>>>(x, ldx, cmptr, cmem, n_clusters, n_clusters_dev, C, y, ldy, argmax);
  }
  DRGNN_CHECK_LAUNCH("maxpool_fwd_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_maxpool_bwd(const float* g, int32_t ldg, const int32_t* argmax, const int32_t* cl,
                                 const float* relu_out, int32_t ld_relu, int32_t n_nodes, const int32_t* n_nodes_dev,
                                 int32_t C, float* dx, int32_t lddx, void* stream) {
  DRGNN_REQUIRE(g && argmax && cl && dx, "maxpool_bwd: NULL pointer");
  DRGNN_REQUIRE(n_nodes >= 0 && C > 0 && ldg >= C && lddx >= C, "maxpool_bwd: bad sizes");
  DRGNN_REQUIRE(!relu_out || ld_relu >= C, "maxpool_bwd: ld_relu too small");
  if (n_nodes == 0) return DRGNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t total = (int64_t)n_nodes * C;
  const int blocks = (int)min_i64((total + 255) / 256, (int64_t)device_info().sms * 16);
  maxpool_bwd_kernel<<<blocks, 256, 0, st>>>(g, ldg, argmax, cl, relu_out, ld_relu, n_nodes, n_nodes_dev, C, dx, lddx);
  DRGNN_CHECK_LAUNCH("maxpool_bwd_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_segment_mean_fwd(const float* x, int32_t ldx, const int32_t* seg_ptr, int32_t B, int32_t C, float* r,
                                      int32_t ldr, void* stream) {
  DRGNN_REQUIRE(x && seg_ptr && r, "segment_mean_fwd: NULL pointer");
  DRGNN_REQUIRE(B >= 0 && C > 0 && ldx >= C && ldr >= C, "segment_mean_fwd: bad sizes");
  if (B == 0) return DRGNN_OK;
  const int threads = C <= 32 ? 32 : (C <= 64 ? 64 : (C <= 128 ? 128 : 256));
  segment_mean_fwd_kernel<<<B, threads, 0, (cudaStream_t)stream>>>(x, ldx, seg_ptr, C, r, ldr);
  DRGNN_CHECK_LAUNCH("segment_mean_fwd_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_segment_mean_bwd(const float* g, int32_t ldg, const int32_t* seg_ptr, int32_t B, int32_t C, float* dx,
                                      int32_t lddx, void* stream) {
  DRGNN_REQUIRE(g && seg_ptr && dx, "segment_mean_bwd: NULL pointer");
  DRGNN_REQUIRE(B >= 0 && C > 0 && ldg >= C && lddx >= C, "segment_mean_bwd: bad sizes");
  if (B == 0) return DRGNN_OK;
  segment_mean_bwd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(g, ldg, seg_ptr, C, dx, lddx);
  DRGNN_CHECK_LAUNCH("segment_mean_bwd_kernel");
  return DRGNN_OK;
}
