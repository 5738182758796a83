// Shared helpers for libdrgnn (sm_100a).  Internal header, not part of the C-ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>

#include "../../include/drgnn.h"

namespace drgnn {

// ---- error plumbing (thread-local message, C-ABI returns an int) ----
inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}
#define DRGNN_CHECK_CUDA(expr)                                                              \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return drgnn::fail(DRGNN_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                               \
  } while (0)
#define DRGNN_CHECK_LAUNCH(name)                                                            \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess)                                                                  \
      return drgnn::fail(DRGNN_ERR_CUDA, "launch of %s failed: %s", name, cudaGetErrorString(_e)); \
  } while (0)
#define DRGNN_REQUIRE(cond, ...)                                   \
  do {                                                             \
    if (!(cond)) return drgnn::fail(DRGNN_ERR_INVALID, __VA_ARGS__); \
  } while (0)

static inline int64_t min_i64(int64_t a, int64_t b) { return a < b ? a : b; }

struct DeviceInfo {
  int sms;
  int smem_optin;
  int device;
};
const DeviceInfo& device_info();

// ---- device helpers ----
__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// inclusive warp scan
__device__ __forceinline__ int warp_scan_incl(int v) {
  const int l = lane_id();
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, v, o);
    if (l >= o) v += t;
  }
  return v;
}

// In-place exclusive scan of a shared-memory int array by the whole CTA.
// `wsum` is a 33-int shared scratch.  Returns the total.  All threads must call.
__device__ inline int block_exclusive_scan(int* a, int len, int* wsum) {
  const int T = blockDim.x, t = threadIdx.x;
  const int chunk = (len + T - 1) / T;
  const int beg = min(t * chunk, len), end = min(beg + chunk, len);
  int s = 0;
  for (int i = beg; i < end; ++i) s += a[i];
  int incl = warp_scan_incl(s);
  if (lane_id() == 31) wsum[warp_id()] = incl;
  __syncthreads();
  if (warp_id() == 0) {
    int nw = T >> 5;
    int v = lane_id() < nw ? wsum[lane_id()] : 0;
    int vi = warp_scan_incl(v);
    wsum[lane_id()] = vi - v;  // exclusive warp offsets
    if (lane_id() == 31) wsum[32] = vi;
  }
  __syncthreads();
  int run = wsum[warp_id()] + incl - s;
  for (int i = beg; i < end; ++i) {
    int v = a[i];
    a[i] = run;
    run += v;
  }
  int total = wsum[32];
  __syncthreads();
  return total;
}

// Adam bias correction 1 - beta^step, the same expression in every optimiser kernel (single GPU, peer
// exchange, in-kernel reduction) so that they stay bit-identical: -expm1(step * log1p(beta - 1)) is
// accurate for beta near 1 and needs no double-precision pow on the critical path.
__device__ __forceinline__ float adam_bias_correction(float beta, float step) {
  return -expm1f(step * log1pf(beta - 1.f));
}

// Word offsets of the arrays inside one graph's structure blob (include/drgnn.h, drgnn_structure_io.blob)
struct BlobLayout {
  int rp0, col0, rp1, col1, cmp0, cmem0, cl0, cmp1, cmem1, cl1, cscp1, cscr1, end;
};
__host__ __device__ inline BlobLayout blob_layout(int n, int m) {
  BlobLayout b;
  int o = DRGNN_BLOB_HEADER;
  b.rp0 = o;   o += n + 1;
  b.col0 = o;  o += m;
  b.rp1 = o;   o += n + 1;
  b.col1 = o;  o += m;
  b.cmp0 = o;  o += n + 1;
  b.cmem0 = o; o += n;
  b.cl0 = o;   o += n;
  b.cmp1 = o;  o += n + 1;
  b.cmem1 = o; o += n;
  b.cl1 = o;   o += n;
  b.cscp1 = o; o += n + 1;
  b.cscr1 = o; o += m;
  b.end = o;
  return b;
}

// CTA-wide small GEMM out of shared memory:  C[m][n] = sum_k At[k * lda + m] * Bm[k * ldb + n]
// (A is held TRANSPOSED so the 8 rows of a thread tile are two 16-byte loads, broadcast inside a
// warp; Bm rows are contiguous in n).  8 (m) x 4 (n) register tile per thread: 3 LDS.128 + 32 FMA per
// k step.  M % 8 == 0 and N % 4 == 0 (pad the buffers), lda % 4 == 0, ldb % 4 == 0, 16-byte aligned
// bases.  out(m, n, value) is called for every element of the padded tile.
template <typename FO>
__device__ __forceinline__ void tile_gemm(const float* __restrict__ At, int lda, const float* __restrict__ Bm, int ldb,
                                          int M, int N, int K, FO out) {
  const int mt = M >> 3, nt = N >> 2;  // M % 8 == 0, N % 4 == 0 (checked on the host)
  for (int item = threadIdx.x; item < mt * nt; item += blockDim.x) {
    const int mg = item / nt, ng = item - mg * nt;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* ap = At + mg * 8;
    const float* bp = Bm + ng * 4;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(ap + k * lda);
      const float4 a1 = *reinterpret_cast<const float4*>(ap + k * lda + 4);
      const float4 b = *reinterpret_cast<const float4*>(bp + k * ldb);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i][0] = fmaf(av[i], b.x, acc[i][0]);
        acc[i][1] = fmaf(av[i], b.y, acc[i][1]);
        acc[i][2] = fmaf(av[i], b.z, acc[i][2]);
        acc[i][3] = fmaf(av[i], b.w, acc[i][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) out(mg * 8 + i, ng * 4 + j, acc[i][j]);
  }
}

}  // namespace drgnn
