// Aggregation: out[i] = act(selfc_i*self[i] + post_i * sum_{p in row i} ew[p]*sscale[col[p]]*src[col[p]] + bias)
//
// Replaces x[col] gather + scatter_sum (deeprank_gnn/ginet.py:57-71), scatter_mean
// (sGAT.py:70-81) and FoutLayer's per-node loop (foutnet.py:71-73).  The same kernels on
// the CSC form compute the backward (A^T g).  No float atomics: a sub-warp of C/4 lanes owns
// one destination row, reads 16 B per lane per neighbour, accumulates in registers and
// issues one coalesced float4 store -> deterministic, summation order = ascending edge id
// (identical to torch_scatter's sequential CPU order).
//
//  * aggregate_rows_kernel : generic CSR rows, source rows come from L1/L2.
//  * aggregate_tiled_kernel: persistent CTAs; the source rows of one graph (tile) are staged
//    in shared memory with cp.async.bulk (TMA 1-D bulk copy, mbarrier complete_tx), double
//    buffered, so every source row is read from HBM exactly once.
#include "common.cuh"

namespace drgnn {

struct AggParams {
  drgnn_aggregate_args a;
  const int32_t* tile_ptr;
  int32_t n_tiles;
  int32_t max_tile_rows;
};

__device__ __forceinline__ float relu_keep_nan(float v) { return v < 0.f ? 0.f : v; }

__device__ __forceinline__ float post_scale(int mode, int deg) {
  if (mode == 1) return 1.f / (float)max(deg, 1);
  if (mode == 2) return 1.f / (float)deg;  // deg == 0 -> inf, inf * 0 = NaN (mean of an empty set)
  return 1.f;
}

// One sub-warp of G lanes per row; lane `sl` owns channels [4*sl, 4*sl+4).
// SrcFn maps a source node id to a pointer to its row (global or shared memory).
template <int G, typename SrcFn>
__device__ __forceinline__ void aggregate_row(const drgnn_aggregate_args& a, int row, int sl, SrcFn src_row) {
  const int C = a.C;
  const bool on = (sl * 4) < C;
  const int s = a.rowptr[row], e = a.rowptr[row + 1];
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float wsum = 0.f;
  int p = s;
  for (; p + 4 <= e; p += 4) {
    int c[4];
    float w[4];
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) c[u] = __ldg(a.col + p + u);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float we = a.ew ? __ldg(a.ew + p + u) : 1.f;
      wsum += we;
      w[u] = a.sscale ? we * __ldg(a.sscale + c[u]) : we;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u)
      v[u] = on ? *reinterpret_cast<const float4*>(src_row(c[u]) + sl * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc.x = fmaf(w[u], v[u].x, acc.x);
      acc.y = fmaf(w[u], v[u].y, acc.y);
      acc.z = fmaf(w[u], v[u].z, acc.z);
      acc.w = fmaf(w[u], v[u].w, acc.w);
    }
  }
  for (; p < e; ++p) {
    const int c = __ldg(a.col + p);
    float we = a.ew ? __ldg(a.ew + p) : 1.f;
    wsum += we;
    const float w = a.sscale ? we * __ldg(a.sscale + c) : we;
    if (on) {
      const float4 v = *reinterpret_cast<const float4*>(src_row(c) + sl * 4);
      acc.x = fmaf(w, v.x, acc.x);
      acc.y = fmaf(w, v.y, acc.y);
      acc.z = fmaf(w, v.z, acc.z);
      acc.w = fmaf(w, v.w, acc.w);
    }
  }
  const float post = post_scale(a.post_mode, e - s);
  acc.x *= post; acc.y *= post; acc.z *= post; acc.w *= post;
  float selfc = 0.f;
  if (a.self_mode == 1) selfc = 1.f;
  else if (a.self_mode == 2) selfc = post * wsum;
  else if (a.self_mode == 3) selfc = __ldg(a.selfc_in + row);
  if (a.self_mode == 2 && a.selfc_out && sl == 0) a.selfc_out[row] = selfc;
  if (a.post_out && sl == 0) a.post_out[row] = post;
  if (!on) return;
  if (a.self_mode != 0 && a.self_src) {
    const float4 sv = *reinterpret_cast<const float4*>(a.self_src + (int64_t)row * a.ld_self + sl * 4);
    if (a.self_out) {
      float4 so = make_float4(selfc * sv.x, selfc * sv.y, selfc * sv.z, selfc * sv.w);
      *reinterpret_cast<float4*>(a.self_out + (int64_t)row * a.ld_self_out + sl * 4) = so;
    } else {
      acc.x = fmaf(selfc, sv.x, acc.x);
      acc.y = fmaf(selfc, sv.y, acc.y);
      acc.z = fmaf(selfc, sv.z, acc.z);
      acc.w = fmaf(selfc, sv.w, acc.w);
    }
  }
  if (a.bias) {
    const float4 b = *reinterpret_cast<const float4*>(a.bias + sl * 4);
    acc.x += b.x; acc.y += b.y; acc.z += b.z; acc.w += b.w;
  }
  if (a.relu) {
    acc.x = relu_keep_nan(acc.x); acc.y = relu_keep_nan(acc.y);
    acc.z = relu_keep_nan(acc.z); acc.w = relu_keep_nan(acc.w);
  }
  *reinterpret_cast<float4*>(a.out + (int64_t)row * a.ld_out + sl * 4) = acc;
}

template <int G>
__global__ void __launch_bounds__(256) aggregate_rows_kernel(const drgnn_aggregate_args a) {
  const int n = a.n_rows_dev ? min(*a.n_rows_dev, a.n_rows) : a.n_rows;
  constexpr int RPW = 32 / G;
  const int lane = lane_id();
  const int sub = lane / G, sl = lane % G;
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  const int wg = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const float* src = a.src;
  const int ld = a.ld_src;
  for (int row = wg * RPW + sub; row < n; row += warps_total * RPW)
    aggregate_row<G>(a, row, sl, [src, ld](int c) { return src + (int64_t)c * ld; });
}

// generic scalar fallback (any C, any alignment): one warp per row, lanes stride over channels
__global__ void __launch_bounds__(256) aggregate_rows_scalar_kernel(const drgnn_aggregate_args a) {
  const int n = a.n_rows_dev ? min(*a.n_rows_dev, a.n_rows) : a.n_rows;
  const int lane = lane_id();
  const int warps_total = (gridDim.x * blockDim.x) >> 5;
  for (int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < n; row += warps_total) {
    const int s = a.rowptr[row], e = a.rowptr[row + 1];
    const float post = post_scale(a.post_mode, e - s);
    float wsum = 0.f;
    for (int p = s; p < e; ++p) wsum += a.ew ? a.ew[p] : 1.f;
    float selfc = 0.f;
    if (a.self_mode == 1) selfc = 1.f;
    else if (a.self_mode == 2) selfc = post * wsum;
    else if (a.self_mode == 3) selfc = a.selfc_in[row];
    if (a.self_mode == 2 && a.selfc_out && lane == 0) a.selfc_out[row] = selfc;
    if (a.post_out && lane == 0) a.post_out[row] = post;
    for (int ch = lane; ch < a.C; ch += 32) {
      float acc = 0.f;
      for (int p = s; p < e; ++p) {
        const int c = a.col[p];
        float w = a.ew ? a.ew[p] : 1.f;
        if (a.sscale) w *= a.sscale[c];
        acc = fmaf(w, a.src[(int64_t)c * a.ld_src + ch], acc);
      }
      acc *= post;
      if (a.self_mode != 0 && a.self_src) {
        const float sv = a.self_src[(int64_t)row * a.ld_self + ch];
        if (a.self_out) a.self_out[(int64_t)row * a.ld_self_out + ch] = selfc * sv;
        else acc = fmaf(selfc, sv, acc);
      }
      if (a.bias) acc += a.bias[ch];
      if (a.relu) acc = relu_keep_nan(acc);
      a.out[(int64_t)row * a.ld_out + ch] = acc;
    }
  }
}

// ---- TMA bulk-copy helpers (cp.async.bulk + mbarrier) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <int G>
__global__ void __launch_bounds__(256) aggregate_tiled_kernel(const AggParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bars[2];
  const drgnn_aggregate_args& a = P.a;
  const int C = a.C;
  const int buf_floats = P.max_tile_rows * C;
  float* bufs = reinterpret_cast<float*>(smem_raw);
  constexpr int RPW = 32 / G;
  const int lane = lane_id(), sub = lane / G, sl = lane % G;
  const int warp = warp_id(), nwarps = blockDim.x >> 5;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // static round-robin over tiles: tile = blockIdx.x + it * gridDim.x
  auto issue = [&](int tile, int b) {
    const int r0 = P.tile_ptr[tile], r1 = P.tile_ptr[tile + 1];
    const uint32_t bytes = (uint32_t)(r1 - r0) * C * 4u;
    if (bytes) {
      mbar_expect_tx(&bars[b], bytes);
      bulk_g2s(bufs + (size_t)b * buf_floats, a.src + (int64_t)r0 * C, bytes, &bars[b]);
    } else {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bars[b])) : "memory");
    }
  };
  int tile = blockIdx.x;
  if (tile < P.n_tiles && threadIdx.x == 0) issue(tile, 0);
  uint32_t phase[2] = {0u, 0u};
  for (int it = 0; tile < P.n_tiles; ++it, tile += gridDim.x) {
    const int b = it & 1;
    const int next = tile + gridDim.x;
    if (next < P.n_tiles && threadIdx.x == 0) issue(next, b ^ 1);  // buffer b^1 was released by the
                                                                   // __syncthreads() closing tile it-1
    mbar_wait(&bars[b], phase[b]);
    phase[b] ^= 1u;
    const int r0 = P.tile_ptr[tile], r1 = P.tile_ptr[tile + 1];
    const float* sbuf = bufs + (size_t)b * buf_floats;
    for (int row = r0 + warp * RPW + sub; row < r1; row += nwarps * RPW)
      aggregate_row<G>(a, row, sl, [sbuf, r0, C](int c) { return sbuf + (size_t)(c - r0) * C; });
    __syncthreads();
  }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static bool vector_ok(const drgnn_aggregate_args& a) {
  if (a.C % 4 != 0 || a.C > 128 || a.C <= 0) return false;
  if (a.ld_src % 4 != 0 || a.ld_out % 4 != 0) return false;
  if (!aligned16(a.src) || !aligned16(a.out)) return false;
  if (a.self_mode != 0 && a.self_src) {
    if (a.ld_self % 4 != 0 || !aligned16(a.self_src)) return false;
    if (a.self_out && (a.ld_self_out % 4 != 0 || !aligned16(a.self_out))) return false;
  }
  if (a.bias && !aligned16(a.bias)) return false;
  return true;
}

static int check_args(const drgnn_aggregate_args* a) {
  DRGNN_REQUIRE(a != nullptr, "aggregate: args is NULL");
  DRGNN_REQUIRE(a->n_rows >= 0 && a->C > 0, "aggregate: bad sizes (n_rows=%d C=%d)", a->n_rows, a->C);
  DRGNN_REQUIRE(a->src && a->out && a->rowptr && a->col, "aggregate: NULL pointer");
  DRGNN_REQUIRE(a->post_mode >= 0 && a->post_mode <= 2, "aggregate: bad post_mode %d", a->post_mode);
  DRGNN_REQUIRE(a->self_mode >= 0 && a->self_mode <= 3, "aggregate: bad self_mode %d", a->self_mode);
  DRGNN_REQUIRE(a->self_mode != 3 || a->selfc_in, "aggregate: self_mode 3 needs selfc_in");
  DRGNN_REQUIRE(a->ld_src >= a->C && a->ld_out >= a->C, "aggregate: leading dimension < C");
  return DRGNN_OK;
}

}  // namespace drgnn

using namespace drgnn;

extern "C" int drgnn_aggregate(const drgnn_aggregate_args* a, void* stream) {
  int rc = check_args(a);
  if (rc) return rc;
  if (a->n_rows == 0) return DRGNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int sms = device_info().sms;
  if (!vector_ok(*a)) {
    int blocks = min((a->n_rows + 7) / 8, sms * 8);
    aggregate_rows_scalar_kernel<<<blocks, 256, 0, st>>>(*a);
    DRGNN_CHECK_LAUNCH("aggregate_rows_scalar_kernel");
    return DRGNN_OK;
  }
  const int lanes = a->C / 4;
  int G = 1;
  while (G < lanes) G <<= 1;
  const int rows_per_block = 8 * (32 / G);
  int blocks = min((a->n_rows + rows_per_block - 1) / rows_per_block, sms * 16);
  switch (G) {
    case 1: aggregate_rows_kernel<1><<<blocks, 256, 0, st>>>(*a); break;
    case 2: aggregate_rows_kernel<2><<<blocks, 256, 0, st>>>(*a); break;
    case 4: aggregate_rows_kernel<4><<<blocks, 256, 0, st>>>(*a); break;
    case 8: aggregate_rows_kernel<8><<<blocks, 256, 0, st>>>(*a); break;
    case 16: aggregate_rows_kernel<16><<<blocks, 256, 0, st>>>(*a); break;
    default: aggregate_rows_kernel<32><<<blocks, 256, 0, st>>>(*a); break;
  }
  DRGNN_CHECK_LAUNCH("aggregate_rows_kernel");
  return DRGNN_OK;
}

template <int G>
static int launch_tiled(const AggParams& P, int blocks, size_t smem, cudaStream_t st) {
  static thread_local size_t configured = 0;
  if (smem > configured) {
    DRGNN_CHECK_CUDA(cudaFuncSetAttribute(aggregate_tiled_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)device_info().smem_optin - 2048));
    configured = device_info().smem_optin - 2048;
  }
  aggregate_tiled_kernel<G><<<blocks, 256, smem, st>>>(P);
  DRGNN_CHECK_LAUNCH("aggregate_tiled_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_aggregate_tiled(const drgnn_aggregate_args* a, const int32_t* tile_ptr, int32_t n_tiles,
                                     int32_t max_tile_rows, void* stream) {
  int rc = check_args(a);
  if (rc) return rc;
  DRGNN_REQUIRE(tile_ptr != nullptr && n_tiles >= 0 && max_tile_rows > 0, "aggregate_tiled: bad tiles");
  DRGNN_REQUIRE(a->n_rows_dev == nullptr, "aggregate_tiled: n_rows_dev is not supported (tiles define the rows)");
  if (n_tiles == 0) return DRGNN_OK;
  if (!vector_ok(*a) || a->ld_src != a->C)
    return fail(DRGNN_ERR_UNSUPPORTED, "aggregate_tiled: needs C %% 4 == 0, C <= 128, ld_src == C, 16-byte alignment");
  const size_t smem = (size_t)2 * max_tile_rows * a->C * sizeof(float);
  if (smem > (size_t)device_info().smem_optin - 2048)
    return fail(DRGNN_ERR_UNSUPPORTED, "aggregate_tiled: tile of %d rows x %d channels does not fit shared memory",
                max_tile_rows, a->C);
  AggParams P;
  P.a = *a;
  P.tile_ptr = tile_ptr;
  P.n_tiles = n_tiles;
  P.max_tile_rows = max_tile_rows;
  const int sms = device_info().sms;
  // CTAs per SM limited by the double buffer; 227 KB usable per SM
  int per_sm = (int)min_i64(8, (size_t)(227 * 1024) / (smem + 1024));
  per_sm = max(per_sm, 1);
  const int blocks = min(n_tiles, sms * per_sm);
  cudaStream_t st = (cudaStream_t)stream;
  const int lanes = a->C / 4;
  int G = 1;
  while (G < lanes) G <<= 1;
  switch (G) {
    case 1: return launch_tiled<1>(P, blocks, smem, st);
    case 2: return launch_tiled<2>(P, blocks, smem, st);
    case 4: return launch_tiled<4>(P, blocks, smem, st);
    case 8: return launch_tiled<8>(P, blocks, smem, st);
    case 16: return launch_tiled<16>(P, blocks, smem, st);
    default: return launch_tiled<32>(P, blocks, smem, st);
  }
}
