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
  const int32_t* tile_ptr;   // [n_tiles+1] row range of every tile
  const int32_t* tile_eptr;  // [n_tiles+1] CSR slot range of every tile (= rowptr[tile_ptr[t]])
  int32_t n_tiles;
  int32_t max_tile_rows;
  int32_t max_tile_edges;
  int32_t stages;            // shared-memory ring depth
};

__host__ __device__ inline int round4(int x) { return (x + 3) & ~3; }
// shared-memory floats (4-byte words) of one staging buffer: x tile | rowptr slice | col slice | ew slice
__host__ __device__ inline int tile_buf_words(int max_rows, int max_edges, int C, bool weights) {
  return max_rows * C + round4(max_rows + 1 + 4) + round4(max_edges + 4) * (weights ? 2 : 1);
}

__device__ __forceinline__ float relu_keep_nan(float v) { return v < 0.f ? 0.f : v; }

__device__ __forceinline__ float post_scale(int mode, int deg) {
  if (mode == 1) return 1.f / (float)max(deg, 1);
  if (mode == 2) return 1.f / (float)deg;  // deg == 0 -> inf, inf * 0 = NaN (mean of an empty set)
  return 1.f;
}

// One sub-warp of G lanes per row; lane `sl` owns channels [4*sl, 4*sl+4).
// SrcFn maps a source node id to a pointer to its row (global or shared memory).
//
// Inner loop (the part that has to run at HBM speed): the G lanes of the sub-warp load G
// consecutive `col` entries (and weights) with ONE coalesced instruction, then broadcast them
// with shuffles and issue up to U independent 16-byte row loads before accumulating, in
// ascending edge order.  WEIGHTED / EPI are compile-time so the plain sum (GINet forward and
// backward, the headline case) carries no per-edge branch, no weight traffic and no epilogue.
//
// rp / cl / ew point at the CSR arrays such that rp[row], cl[p], ew[p] are valid for the absolute
// row / slot ids used here: the global arrays themselves, or shared-memory copies of one tile's
// slices (pre-offset), in which case nothing in the loop touches global memory but `sscale`.
// Everything after the neighbour sum: post scale (sum / mean / mean-or-NaN), self term, bias, ReLU.
template <bool WEIGHTED, bool EPI>
__device__ __forceinline__ void row_epilogue(const drgnn_aggregate_args& a, int row, int sl, bool on, float4 acc,
                                             float wsum, int deg) {
  if (!EPI) {
    if (on) *reinterpret_cast<float4*>(a.out + (int64_t)row * a.ld_out + sl * 4) = acc;
    return;
  }
  const float post = post_scale(a.post_mode, deg);
  acc.x *= post; acc.y *= post; acc.z *= post; acc.w *= post;
  float selfc = 0.f;
  if (a.self_mode == 1) selfc = 1.f;
  else if (a.self_mode == 2) {
    if (!WEIGHTED) wsum = (float)deg;
    selfc = post * wsum;
  } else if (a.self_mode == 3) selfc = __ldg(a.selfc_in + row);
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

// The per-row work is split in three stages so that the global-memory kernel can software-pipeline
// them across row groups (stage 1 of group k+2 and stage 2 of group k+1 are in flight while
// stage 3 of group k gathers): (1) row bounds, (2) the first G col / weight entries of the row,
// loaded cooperatively by the G lanes of the sub-warp, (3) gather + accumulate + epilogue.
// All stages are called by ALL 32 lanes of the warp with warp-uniform control flow, so every
// shuffle uses the full mask and compiles to a single SHFL; `valid` = this sub-warp owns a row.
struct RowSpan {
  int s, e;
};
struct RowCols {
  int c;
  float w;     // weight of the lane's slot (ew * sscale), 0 past the row end
  float wraw;  // ew alone (self_mode 2 needs sum of ew)
};

__device__ __forceinline__ RowSpan load_span(const int32_t* __restrict__ rp, int row, bool valid) {
  RowSpan r;
  r.s = 0;
  r.e = 0;
  if (valid) {
    r.s = rp[row];
    r.e = rp[row + 1];
  }
  return r;
}

template <bool WEIGHTED>
__device__ __forceinline__ RowCols load_cols(const drgnn_aggregate_args& a, const int32_t* __restrict__ cl,
                                             const float* __restrict__ ew, int my, int e) {
  RowCols r;
  r.c = 0;
  r.w = 0.f;
  r.wraw = 0.f;
  if (my < e) {
    r.c = cl[my];
    if (WEIGHTED) {
      r.wraw = ew ? ew[my] : 1.f;
      r.w = a.sscale ? r.wraw * __ldg(a.sscale + r.c) : r.wraw;
    }
  }
  return r;
}

template <int G, bool WEIGHTED, bool EPI, typename SrcFn>
__device__ __forceinline__ void aggregate_row(const drgnn_aggregate_args& a, const int32_t* __restrict__ cl,
                                              const float* __restrict__ ew, int row, bool valid, RowSpan sp,
                                              RowCols first, int sl, bool on, SrcFn src_row) {
  constexpr int U = G < 8 ? G : 8;
  constexpr unsigned FULL = 0xffffffffu;
  const int s = sp.s, e = sp.e;
  const int maxdeg = __reduce_max_sync(FULL, e - s);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  float wsum = 0.f;
  for (int base = 0; base < maxdeg; base += G) {
    RowCols rc = first;
    if (base > 0) rc = load_cols<WEIGHTED>(a, cl, ew, s + base + sl, e);  // rows longer than G (rare)
    if (WEIGHTED) wsum += rc.wraw;
    const int cnt = e - s - base;            // edges left in this sub-warp's row (may be <= 0)
    const int wcnt = min(G, maxdeg - base);  // warp-uniform trip count
#pragma unroll
    for (int j0 = 0; j0 < G; j0 += U) {
      if (j0 >= wcnt) break;
      float4 v[U];
      float wj[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int cj = __shfl_sync(FULL, rc.c, j0 + u, G);
        if (WEIGHTED) wj[u] = __shfl_sync(FULL, rc.w, j0 + u, G);
        if (on && (j0 + u) < cnt) v[u] = *reinterpret_cast<const float4*>(src_row(cj) + sl * 4);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (on && (j0 + u) < cnt) {
          if (WEIGHTED) {
            acc.x = fmaf(wj[u], v[u].x, acc.x);
            acc.y = fmaf(wj[u], v[u].y, acc.y);
            acc.z = fmaf(wj[u], v[u].z, acc.z);
            acc.w = fmaf(wj[u], v[u].w, acc.w);
          } else {
            acc.x += v[u].x;
            acc.y += v[u].y;
            acc.z += v[u].z;
            acc.w += v[u].w;
          }
        }
      }
    }
  }
  if (EPI && WEIGHTED && a.self_mode == 2) {
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) wsum += __shfl_xor_sync(FULL, wsum, o, G);
  }
  if (!valid) return;
  row_epilogue<WEIGHTED, EPI>(a, row, sl, on, acc, wsum, e - s);
}

// Each CTA walks contiguous chunks of `chunk_rows` rows (graphs are contiguous row ranges, so the
// neighbour rows a chunk gathers stay in this SM's L1), chunks are dealt round-robin to CTAs.
// (A software-pipelined variant that prefetched the row bounds / col entries of the next row
// groups measured slower on B200 - 0.69 ms vs 0.50 ms on the 6.5 M-node stream - and was dropped.)
template <int G, bool WEIGHTED, bool EPI>
__global__ void __launch_bounds__(256) aggregate_rows_kernel(const drgnn_aggregate_args a, int chunk_rows) {
  const int n = a.n_rows_dev ? min(*a.n_rows_dev, a.n_rows) : a.n_rows;
  constexpr int RPW = 32 / G;
  const int lane = lane_id();
  const int sub = lane / G, sl = lane % G;
  const bool on = (sl * 4) < a.C;
  const int warp = warp_id(), nwarps = blockDim.x >> 5;
  const float* __restrict__ src = a.src;
  const int ld = a.ld_src;
  const int32_t* __restrict__ rp = a.rowptr;
  const int32_t* __restrict__ cl = a.col;
  const float* __restrict__ ew = a.ew;
  auto src_row = [src, ld](int c) { return src + (int64_t)c * ld; };
  for (int c0 = blockIdx.x * chunk_rows; c0 < n; c0 += gridDim.x * chunk_rows) {
    const int c1 = min(c0 + chunk_rows, n);
    for (int rbase = c0 + warp * RPW; rbase < c1; rbase += nwarps * RPW) {  // warp-uniform
      const int row = rbase + sub;
      const bool valid = row < c1;
      const RowSpan sp = load_span(rp, row, valid);
      const RowCols rc = load_cols<WEIGHTED>(a, cl, ew, sp.s + sl, sp.e);
      aggregate_row<G, WEIGHTED, EPI>(a, cl, ew, row, valid, sp, rc, sl, on, src_row);
    }
  }
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

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds32(uint32_t addr) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ float ldsf32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Warp-specialised multi-stage TMA pipeline, ONE persistent CTA per SM.
//   warp 0 (one elected lane) = producer: streams whole tiles (graphs) - feature rows, rowptr slice,
//     col slice, weight slice - into a ring of `stages` shared-memory buffers with bulk async copies
//     (cp.async.bulk) that complete on the stage's FULL mbarrier; it only waits for the stage's EMPTY
//     mbarrier, so up to `stages` tiles (~200 KB per SM) are in flight: enough bytes outstanding to
//     cover HBM latency at full bandwidth, which neither a 2-buffer scheme nor the L1-resident
//     global-memory kernel achieves.
//   warps 1..kConsumers = consumers: wait on FULL, process their share of the tile's rows entirely
//     out of shared memory (explicit ld.shared with 32-bit addresses), arrive on EMPTY.  No CTA-wide
//     barrier in the loop; row groups are dealt to warps from a per-tile rotating offset.
#ifndef DRGNN_CONSUMERS
#define DRGNN_CONSUMERS 31
#endif
#ifndef DRGNN_MAX_STAGES
#define DRGNN_MAX_STAGES 8
#endif
static constexpr int kConsumers = DRGNN_CONSUMERS;
static constexpr int kTiledThreads = 32 * (1 + kConsumers);
static constexpr int kMaxStages = DRGNN_MAX_STAGES;

template <int G, bool WEIGHTED, bool EPI>
__global__ void __launch_bounds__(kTiledThreads, 1) aggregate_tiled_kernel(const AggParams P) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  __shared__ int tinfo[kMaxStages][4];  // per stage: r0, r1, aligned rowptr start, aligned slot start
  const drgnn_aggregate_args& a = P.a;
  const int C = a.C;
  const int S = P.stages;
  const bool has_w = WEIGHTED && a.ew != nullptr;
  const int rp_words = round4(P.max_tile_rows + 1 + 4), cl_words = round4(P.max_tile_edges + 4);
  const int buf_words = tile_buf_words(P.max_tile_rows, P.max_tile_edges, C, has_w);
  float* bufs = reinterpret_cast<float*>(smem_raw);
  constexpr int RPW = 32 / G;
  const int lane = lane_id(), warp = warp_id();

  if (threadIdx.x == 0) {
    for (int i = 0; i < S; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], kConsumers);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == 0) {
    // ---------------- producer ----------------
    if (lane != 0) return;
    int st = 0;
    uint32_t fill = 0;  // completed ring revolutions
    for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
      if (fill >= 1) mbar_wait(&empty_bar[st], (fill - 1) & 1u);
      // slices start at 16-byte aligned element ids (arrays are readable up to the next multiple of
      // 4 elements past their end, see drgnn.h)
      const int r0 = __ldg(P.tile_ptr + tile), r1 = __ldg(P.tile_ptr + tile + 1);
      const int e0 = __ldg(P.tile_eptr + tile), e1 = __ldg(P.tile_eptr + tile + 1);
      const int ra = r0 & ~3, rb = round4(r1 + 1);
      const int ea = e0 & ~3, eb = round4(e1);
      float* xs = bufs + (size_t)st * buf_words;
      int32_t* rps = reinterpret_cast<int32_t*>(xs + P.max_tile_rows * C);
      int32_t* cls = rps + rp_words;
      float* ews = reinterpret_cast<float*>(cls + cl_words);
      const uint32_t bx = (uint32_t)(r1 - r0) * C * 4u, br = (uint32_t)(rb - ra) * 4u, be = (uint32_t)(eb - ea) * 4u;
      tinfo[st][0] = r0; tinfo[st][1] = r1; tinfo[st][2] = ra; tinfo[st][3] = ea;
      mbar_expect_tx(&full_bar[st], bx + br + be * (has_w ? 2u : 1u));
      if (bx) bulk_g2s(xs, a.src + (int64_t)r0 * C, bx, &full_bar[st]);
      bulk_g2s(rps, a.rowptr + ra, br, &full_bar[st]);
      if (be) {
        bulk_g2s(cls, a.col + ea, be, &full_bar[st]);
        if (has_w) bulk_g2s(ews, a.ew + ea, be, &full_bar[st]);
      }
      if (++st == S) { st = 0; ++fill; }
    }
    return;
  }

  // ---------------- consumers ----------------
  const int cw = warp - 1;
  const int sub = lane / G, sl = lane % G;
  const bool on = (sl * 4) < C;
  int st = 0, rot = 0;
  uint32_t phase = 0;
  for (int tile = blockIdx.x; tile < P.n_tiles; tile += gridDim.x) {
    mbar_wait(&full_bar[st], phase);
    const int r0 = tinfo[st][0], r1 = tinfo[st][1], ra = tinfo[st][2], ea = tinfo[st][3];
    const uint32_t xs_a = smem_u32(bufs + (size_t)st * buf_words);
    const uint32_t rp_a = xs_a + (uint32_t)(P.max_tile_rows * C) * 4u;  // rowptr slice, element `ra` first
    const uint32_t cl_a = rp_a + (uint32_t)rp_words * 4u;                // col slice, element `ea` first
    const uint32_t ew_a = cl_a + (uint32_t)cl_words * 4u;
    const uint32_t lane_a = xs_a + (uint32_t)sl * 16u;
    const int ngroups = (r1 - r0 + RPW - 1) / RPW;
    int g0 = cw - rot;
    if (g0 < 0) g0 += kConsumers;
    for (int g = g0; g < ngroups; g += kConsumers) {
      const int row = r0 + g * RPW + sub;
      if (row < r1) {
        const int s = lds32(rp_a + (uint32_t)(row - ra) * 4u), e = lds32(rp_a + (uint32_t)(row + 1 - ra) * 4u);
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float wsum = 0.f;
        int p = s;
        for (; p + 4 <= e; p += 4) {
          int c[4];
          float w[4];
          float4 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) c[u] = lds32(cl_a + (uint32_t)(p + u - ea) * 4u);
          if (WEIGHTED) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const float we = has_w ? ldsf32(ew_a + (uint32_t)(p + u - ea) * 4u) : 1.f;
              wsum += we;
              w[u] = a.sscale ? we * __ldg(a.sscale + c[u]) : we;
            }
          }
          if (on) {
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = lds128(lane_a + (uint32_t)((c[u] - r0) * C) * 4u);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              if (WEIGHTED) {
                acc.x = fmaf(w[u], v[u].x, acc.x); acc.y = fmaf(w[u], v[u].y, acc.y);
                acc.z = fmaf(w[u], v[u].z, acc.z); acc.w = fmaf(w[u], v[u].w, acc.w);
              } else {
                acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
              }
            }
          }
        }
        for (; p < e; ++p) {
          const int c = lds32(cl_a + (uint32_t)(p - ea) * 4u);
          float w = 1.f;
          if (WEIGHTED) {
            const float we = has_w ? ldsf32(ew_a + (uint32_t)(p - ea) * 4u) : 1.f;
            wsum += we;
            w = a.sscale ? we * __ldg(a.sscale + c) : we;
          }
          if (on) {
            const float4 v = lds128(lane_a + (uint32_t)((c - r0) * C) * 4u);
            if (WEIGHTED) {
              acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y);
              acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
            } else {
              acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
          }
        }
        row_epilogue<WEIGHTED, EPI>(a, row, sl, on, acc, wsum, e - s);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[st]);  // this warp is done reading the stage
    if (++st == S) { st = 0; phase ^= 1u; }
    if (++rot == kConsumers) rot = 0;
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

static inline bool is_weighted(const drgnn_aggregate_args& a) { return a.ew != nullptr || a.sscale != nullptr; }
static inline bool has_epilogue(const drgnn_aggregate_args& a) {
  return a.post_mode != 0 || a.self_mode != 0 || a.bias != nullptr || a.relu != 0 || a.post_out != nullptr;
}

template <int G>
static void launch_rows(const drgnn_aggregate_args& a, int blocks, int chunk_rows, cudaStream_t st) {
  const bool w = is_weighted(a), e = has_epilogue(a);
  if (w && e) aggregate_rows_kernel<G, true, true><<<blocks, 256, 0, st>>>(a, chunk_rows);
  else if (w) aggregate_rows_kernel<G, true, false><<<blocks, 256, 0, st>>>(a, chunk_rows);
  else if (e) aggregate_rows_kernel<G, false, true><<<blocks, 256, 0, st>>>(a, chunk_rows);
  else aggregate_rows_kernel<G, false, false><<<blocks, 256, 0, st>>>(a, chunk_rows);
}

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
  const int rows_per_pass = 8 * (32 / G);  // rows one CTA covers per pass (8 warps)
  // small launches: one pass per CTA (spread over all SMs, latency); large launches: up to 8 passes
  // of contiguous rows per CTA visit (neighbour rows stay in that SM's L1)
  int passes = (int)min_i64(8, (int64_t)a->n_rows / ((int64_t)rows_per_pass * sms * 8));
  if (passes < 1) passes = 1;
  const int chunk_rows = rows_per_pass * passes;
  int blocks = (int)min_i64(((int64_t)a->n_rows + chunk_rows - 1) / chunk_rows, (int64_t)sms * 8);
  switch (G) {
    case 1: launch_rows<1>(*a, blocks, chunk_rows, st); break;
    case 2: launch_rows<2>(*a, blocks, chunk_rows, st); break;
    case 4: launch_rows<4>(*a, blocks, chunk_rows, st); break;
    case 8: launch_rows<8>(*a, blocks, chunk_rows, st); break;
    case 16: launch_rows<16>(*a, blocks, chunk_rows, st); break;
    default: launch_rows<32>(*a, blocks, chunk_rows, st); break;
  }
  DRGNN_CHECK_LAUNCH("aggregate_rows_kernel");
  return DRGNN_OK;
}

template <int G, bool W, bool E>
static int launch_tiled_v(const AggParams& P, int blocks, size_t smem, cudaStream_t st) {
  static thread_local size_t configured = 0;
  if (smem > configured) {
    DRGNN_CHECK_CUDA(cudaFuncSetAttribute(aggregate_tiled_kernel<G, W, E>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)device_info().smem_optin - 2048));
    configured = device_info().smem_optin - 2048;
  }
  aggregate_tiled_kernel<G, W, E><<<blocks, kTiledThreads, smem, st>>>(P);
  DRGNN_CHECK_LAUNCH("aggregate_tiled_kernel");
  return DRGNN_OK;
}

template <int G>
static int launch_tiled(const AggParams& P, int blocks, size_t smem, cudaStream_t st) {
  const bool w = is_weighted(P.a), e = has_epilogue(P.a);
  if (w && e) return launch_tiled_v<G, true, true>(P, blocks, smem, st);
  if (w) return launch_tiled_v<G, true, false>(P, blocks, smem, st);
  if (e) return launch_tiled_v<G, false, true>(P, blocks, smem, st);
  return launch_tiled_v<G, false, false>(P, blocks, smem, st);
}

extern "C" int drgnn_aggregate_tiled(const drgnn_aggregate_args* a, const int32_t* tile_ptr, const int32_t* tile_eptr,
                                     int32_t n_tiles, int32_t max_tile_rows, int32_t max_tile_edges, void* stream) {
  int rc = check_args(a);
  if (rc) return rc;
  DRGNN_REQUIRE(tile_ptr != nullptr && tile_eptr != nullptr && n_tiles >= 0 && max_tile_rows > 0 && max_tile_edges >= 0,
                "aggregate_tiled: bad tiles");
  DRGNN_REQUIRE(a->n_rows_dev == nullptr, "aggregate_tiled: n_rows_dev is not supported (tiles define the rows)");
  if (n_tiles == 0) return DRGNN_OK;
  if (!vector_ok(*a) || a->ld_src != a->C || !aligned16(a->rowptr) || !aligned16(a->col) || (a->ew && !aligned16(a->ew)))
    return fail(DRGNN_ERR_UNSUPPORTED, "aggregate_tiled: needs C %% 4 == 0, C <= 128, ld_src == C, 16-byte alignment");
  const size_t buf_bytes = (size_t)4 * tile_buf_words(max_tile_rows, max_tile_edges, a->C, a->ew != nullptr);
  const size_t budget = (size_t)device_info().smem_optin - 2048;
  int stages = (int)min_i64(kMaxStages, (int64_t)(budget / buf_bytes));
  if (stages < 2)
    return fail(DRGNN_ERR_UNSUPPORTED, "aggregate_tiled: tile of %d rows x %d channels / %d edges does not fit shared memory twice",
                max_tile_rows, a->C, max_tile_edges);
  const int sms = device_info().sms;
  const int blocks = min(n_tiles, sms);  // one persistent CTA per SM
  stages = (int)min_i64(stages, ((int64_t)n_tiles + blocks - 1) / blocks + 1);
  if (stages < 2) stages = 2;
  const size_t smem = buf_bytes * stages;
  AggParams P;
  P.a = *a;
  P.tile_ptr = tile_ptr;
  P.tile_eptr = tile_eptr;
  P.n_tiles = n_tiles;
  P.max_tile_rows = max_tile_rows;
  P.max_tile_edges = max_tile_edges;
  P.stages = stages;
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
