// Dense per-node transform and its weight gradient.
//
// Replaces the nn.Linear / torch.mm calls of the reference layers
// (deeprank_gnn/ginet.py:57-58,137-139, sGAT.py:73,134-135, foutnet.py:62,65,121-122).  The
// reference applies the conv transforms to E gathered rows; here they run on N node rows
// (aggregate-then-transform), optionally on several weight groups at once (GINet's two
// branches are one launch with groups = 2).
//
//   linear_fma_kernel    : Y = act(X W + b) (* mask), 64 x 64 output tile per CTA, 4 x 4
//                          register micro-tiles, fp32 FMA (bit-stable, parity reference).
//   linear_tf32x3_kernel : same contract on the tensor cores: mma.sync.m16n8k8 TF32 with the
//                          3-product error compensation (hi*hi + hi*lo + lo*hi), fp32 accumulate.
//   wgrad_partial_kernel / wgrad_reduce_kernel : dW = G^T X, db = sum_r G as a two-phase
//                          fixed-order reduction over row chunks (deterministic, no atomics).
#include "common.cuh"

namespace drgnn {

static constexpr int LT_R = 64;  // rows per CTA tile
static constexpr int LT_C = 64;  // output columns per CTA tile
static constexpr int LT_K = 32;  // reduction chunk

__device__ __forceinline__ int live_rows(int rows, const int32_t* rows_dev) {
  return rows_dev ? min(*rows_dev, rows) : rows;
}

// Stage one [LT_K x LT_C] chunk of the group's weight as Ws[k][c] (zero padded).
__device__ __forceinline__ void load_w_chunk(const drgnn_linear_args& a, int g, int k0, int o0, float (*Ws)[LT_C]) {
  for (int idx = threadIdx.x; idx < LT_K * LT_C; idx += blockDim.x) {
    const int kk = idx / LT_C, c = idx % LT_C;
    const int k = k0 + kk, o = o0 + c;
    float v = 0.f;
    if (k < a.Fin && o < a.Fout) {
      v = a.w_layout == 0 ? __ldg(a.W + ((int64_t)g * a.Fout + o) * a.Fin + k)
                          : __ldg(a.W + ((int64_t)g * a.Fin + k) * a.Fout + o);
    }
    Ws[kk][c] = v;
  }
}

__device__ __forceinline__ float finish(const drgnn_linear_args& a, float v, int64_t r, int o) {
  if (a.bias) v += __ldg(a.bias + o);
  if (a.relu) v = v < 0.f ? 0.f : v;  // keeps NaN like torch.relu
  if (a.out_mask) v = (a.out_mask[r * a.ld_mask + o] > 0.f) ? v * a.mask_scale : 0.f;
  return v;
}

// 16 * cgt compute threads (cgt = 4-column groups of the tile, <= 16); blockDim.x >= 16 * cgt, the
// surplus threads only help staging the tiles (narrow outputs such as fc2 would otherwise stage a
// 32 x 64 weight chunk with 16 threads); grid = (row tiles, groups * col tiles)
__global__ void __launch_bounds__(256) linear_fma_kernel(const drgnn_linear_args a, int col_tiles, int cgt) {
  __shared__ float Xs[LT_R][LT_K + 1];
  __shared__ __align__(16) float Ws[LT_K][LT_C];
  const int rows = live_rows(a.rows, a.rows_dev);
  const int r0 = blockIdx.x * LT_R;
  if (r0 >= rows) return;
  const int g = blockIdx.y / col_tiles, o0 = (blockIdx.y % col_tiles) * LT_C;
  const bool worker = threadIdx.x < 16 * cgt;
  const int rg = worker ? threadIdx.x / cgt : 0, cg = worker ? threadIdx.x % cgt : 0;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.Fin; k0 += LT_K) {
    for (int idx = threadIdx.x; idx < LT_R * LT_K; idx += blockDim.x) {
      const int rr = idx / LT_K, kk = idx % LT_K;
      const int r = r0 + rr, k = k0 + kk;
      Xs[rr][kk] = (r < rows && k < a.Fin) ? a.X[(int64_t)r * a.ldx + (int64_t)g * a.Fin + k] : 0.f;
    }
    load_w_chunk(a, g, k0, o0, Ws);
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < LT_K; ++kk) {
      const float4 w = *reinterpret_cast<const float4*>(&Ws[kk][cg * 4]);
      float x[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = Xs[rg * 4 + i][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(x[i], w.x, acc[i][0]);
        acc[i][1] = fmaf(x[i], w.y, acc[i][1]);
        acc[i][2] = fmaf(x[i], w.z, acc[i][2]);
        acc[i][3] = fmaf(x[i], w.w, acc[i][3]);
      }
    }
    __syncthreads();
  }
  if (!worker) return;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = r0 + rg * 4 + i;
    if (r >= rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int oc = o0 + cg * 4 + j;
      if (oc >= a.Fout) continue;
      const int o = g * a.Fout + oc;
      a.Y[(int64_t)r * a.ldy + o] = finish(a, acc[i][j], r, o);
    }
  }
}

// ---------------------------------------------------------------------------------------
// Tensor-core path: 3xTF32.  One warp owns a 16-row x 64-col slab of the CTA tile (4 warps =
// 64 rows); per k-chunk of 8 it issues, for each of the 8 n-tiles, three m16n8k8 MMAs
// (a_hi*b_hi, a_hi*b_lo, a_lo*b_hi) so the product keeps ~fp32 accuracy (error ~2^-21).
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = f2tf32(x);
  lo = f2tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(128) linear_tf32x3_kernel(const drgnn_linear_args a, int col_tiles) {
  __shared__ float Xs[LT_R][LT_K + 4];               // stride 36: conflict-free fragment reads
  __shared__ __align__(16) float Ws[LT_K][LT_C + 8];  // stride 72: conflict-free fragment reads
  const int rows = live_rows(a.rows, a.rows_dev);
  const int r0 = blockIdx.x * LT_R;
  if (r0 >= rows) return;
  const int g = blockIdx.y / col_tiles, o0 = (blockIdx.y % col_tiles) * LT_C;
  const int warp = warp_id(), lane = lane_id();
  const int gid = lane >> 2, tig = lane & 3;
  const int ntiles = min(8, (a.Fout - o0 + 7) / 8);
  float acc[8][4];
#pragma unroll
  for (int n = 0; n < 8; ++n)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[n][j] = 0.f;

  for (int k0 = 0; k0 < a.Fin; k0 += LT_K) {
    for (int idx = threadIdx.x; idx < LT_R * LT_K; idx += blockDim.x) {
      const int rr = idx / LT_K, kk = idx % LT_K;
      const int r = r0 + rr, k = k0 + kk;
      Xs[rr][kk] = (r < rows && k < a.Fin) ? a.X[(int64_t)r * a.ldx + (int64_t)g * a.Fin + k] : 0.f;
    }
    for (int idx = threadIdx.x; idx < LT_K * LT_C; idx += blockDim.x) {
      const int kk = idx / LT_C, c = idx % LT_C;
      const int k = k0 + kk, o = o0 + c;
      float v = 0.f;
      if (k < a.Fin && o < a.Fout)
        v = a.w_layout == 0 ? __ldg(a.W + ((int64_t)g * a.Fout + o) * a.Fin + k)
                            : __ldg(a.W + ((int64_t)g * a.Fin + k) * a.Fout + o);
      Ws[kk][c] = v;
    }
    __syncthreads();
    const int kmax = min(LT_K, a.Fin - k0);
    for (int ks = 0; ks < kmax; ks += 8) {
      // A fragment (16x8, row): a0=(gid, tig) a1=(gid+8, tig) a2=(gid, tig+4) a3=(gid+8, tig+4)
      uint32_t ahi[4], alo[4];
      const int ar = warp * 16 + gid;
      split_tf32(Xs[ar][ks + tig], ahi[0], alo[0]);
      split_tf32(Xs[ar + 8][ks + tig], ahi[1], alo[1]);
      split_tf32(Xs[ar][ks + tig + 4], ahi[2], alo[2]);
      split_tf32(Xs[ar + 8][ks + tig + 4], ahi[3], alo[3]);
#pragma unroll
      for (int n = 0; n < 8; ++n) {
        if (n < ntiles) {
          // B fragment (8x8, col): b0=(k=tig, n=gid) b1=(k=tig+4, n=gid)
          uint32_t bhi[2], blo[2];
          split_tf32(Ws[ks + tig][n * 8 + gid], bhi[0], blo[0]);
          split_tf32(Ws[ks + tig + 4][n * 8 + gid], bhi[1], blo[1]);
          mma_tf32(acc[n], alo, bhi);
          mma_tf32(acc[n], ahi, blo);
          mma_tf32(acc[n], ahi, bhi);
        }
      }
    }
    __syncthreads();
  }
  // C fragment: c0=(gid, 2*tig) c1=(gid, 2*tig+1) c2=(gid+8, 2*tig) c3=(gid+8, 2*tig+1)
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    if (n >= ntiles) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + warp * 16 + gid + ((j & 2) ? 8 : 0);
      const int oc = o0 + n * 8 + 2 * tig + (j & 1);
      if (r < rows && oc < a.Fout) {
        const int o = g * a.Fout + oc;
        a.Y[(int64_t)r * a.ldy + o] = finish(a, acc[n][j], r, o);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// Weight gradient
// ---------------------------------------------------------------------------------------
static constexpr int WG_O = 64, WG_K = 64, WG_R = 32;

static inline int wgrad_chunks(int rows) {
  int c = (rows + 127) / 128;
  const int cap = 2 * device_info().sms;
  if (c > cap) c = cap;
  if (c < 1) c = 1;
  return c;
}

// grid = (n_chunks, groups * o_tiles * k_tiles); thread (to, tk) owns a 4 x 4 block of dW
__global__ void __launch_bounds__(256) wgrad_partial_kernel(const drgnn_linear_wgrad_args a, int n_chunks, int o_tiles,
                                                            int k_tiles) {
  __shared__ __align__(16) float Gs[WG_R][WG_O];
  __shared__ __align__(16) float Xs[WG_R][WG_K];
  const int rows = live_rows(a.rows, a.rows_dev);
  int rpc = (rows + n_chunks - 1) / n_chunks;
  rpc = ((rpc + WG_R - 1) / WG_R) * WG_R;
  const int rbeg = min(blockIdx.x * rpc, rows), rend = min(rbeg + rpc, rows);
  int y = blockIdx.y;
  const int kt = y % k_tiles;
  y /= k_tiles;
  const int ot = y % o_tiles;
  const int g = y / o_tiles;
  const int o0 = ot * WG_O, k0 = kt * WG_K;
  const int to = threadIdx.x >> 4, tk = threadIdx.x & 15;
  float acc[4][4], bsum[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    bsum[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  }
  for (int rb = rbeg; rb < rend; rb += WG_R) {
    for (int idx = threadIdx.x; idx < WG_R * WG_O; idx += blockDim.x) {
      const int rr = idx / WG_O, c = idx % WG_O;
      const int r = rb + rr;
      Gs[rr][c] = (r < rend && o0 + c < a.Fout) ? a.G[(int64_t)r * a.ldg + (int64_t)g * a.Fout + o0 + c] : 0.f;
      float xv = (r < rend && k0 + c < a.Fin) ? a.X[(int64_t)r * a.ldx + (int64_t)g * a.Fin + k0 + c] : 0.f;
      // A NaN input row (FoutLayer on a node without neighbour, foutnet.py:73) yields a NaN output row,
      // whose gradient is exactly 0 (ReLU mask / max-pool never select NaN): drop it instead of 0 * NaN.
      Xs[rr][c] = (xv == xv) ? xv : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < WG_R; ++rr) {
      const float4 gv = *reinterpret_cast<const float4*>(&Gs[rr][to * 4]);
      const float4 xv = *reinterpret_cast<const float4*>(&Xs[rr][tk * 4]);
      const float gg[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] = fmaf(gg[i], xv.x, acc[i][0]);
        acc[i][1] = fmaf(gg[i], xv.y, acc[i][1]);
        acc[i][2] = fmaf(gg[i], xv.z, acc[i][2]);
        acc[i][3] = fmaf(gg[i], xv.w, acc[i][3]);
        bsum[i] += gg[i];
      }
    }
    __syncthreads();
  }
  const int64_t nW = (int64_t)a.groups * a.Fout * a.Fin;
  float* part = a.work + (int64_t)blockIdx.x * (nW + (int64_t)a.groups * a.Fout);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int o = o0 + to * 4 + i;
    if (o >= a.Fout) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + tk * 4 + j;
      if (k >= a.Fin) continue;
      const int64_t e = a.w_layout == 0 ? ((int64_t)g * a.Fout + o) * a.Fin + k : ((int64_t)g * a.Fin + k) * a.Fout + o;
      part[e] = acc[i][j];
    }
    if (kt == 0 && tk == 0) part[nW + (int64_t)g * a.Fout + o] = bsum[i];
  }
}

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const drgnn_linear_wgrad_args a, int n_chunks) {
  const int64_t nW = (int64_t)a.groups * a.Fout * a.Fin;
  const int64_t stride = nW + (int64_t)a.groups * a.Fout;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= stride) return;
  float s = 0.f;
  for (int c = 0; c < n_chunks; ++c) s += a.work[(int64_t)c * stride + e];
  if (e < nW) {
    a.dW[e] = a.accumulate ? a.dW[e] + s : s;
  } else if (a.dbias) {
    a.dbias[e - nW] = a.accumulate ? a.dbias[e - nW] + s : s;
  }
}

// ---------------------------------------------------------------------------------------
// Weight gradient, small-matrix fast path (every conv layer of the reference networks):
// per group Fout <= 64 and Fout * ceil(Fin / 32) <= 64 accumulators per lane.
//   * a warp owns a contiguous slice of rows; lane l owns column l (and l + 32) of X and keeps the
//     whole Fout x {1,2} strip of dW in registers; the G row is loaded once, coalesced, and its
//     entries are broadcast with shuffles: Fout SHFL + Fout FMA per row, no shared memory;
//   * the 8 warp strips of a CTA are summed in shared memory in warp order, the CTA partial goes
//     to scratch, and the LAST CTA to arrive (ticket counter) sums the partials in CTA order and
//     writes dW / dbias: one launch, fixed summation order (deterministic), counter self-resets.
// ---------------------------------------------------------------------------------------
static constexpr int WGF_MAX_CTAS = 64;

template <int FOUT, int NC, int NT>
__global__ void __launch_bounds__(NT) wgrad_warp_kernel(const drgnn_linear_wgrad_args a, float* scratch, unsigned* ticket) {
  extern __shared__ float wsm[];  // [NT / 32 warps][E]
  constexpr int NW = NT / 32;
  __shared__ bool is_last;
  const int Fin = a.Fin, Fout = a.Fout, groups = a.groups;
  const int nW = groups * Fout * Fin, E = nW + groups * Fout;
  const int rows = live_rows(a.rows, a.rows_dev);
  const int lane = lane_id(), warp = warp_id();
  const int nwg = gridDim.x * NW, wg = blockIdx.x * NW + warp;
  const int rpw = (rows + nwg - 1) / nwg;
  const int rbeg = min(wg * rpw, rows), rend = min(rbeg + rpw, rows);
  float* mine = wsm + (size_t)warp * E;
  for (int g = 0; g < groups; ++g) {
    float acc[FOUT][NC];
    float bacc[(FOUT + 31) / 32];
#pragma unroll
    for (int o = 0; o < FOUT; ++o)
#pragma unroll
      for (int c = 0; c < NC; ++c) acc[o][c] = 0.f;
#pragma unroll
    for (int q = 0; q < (FOUT + 31) / 32; ++q) bacc[q] = 0.f;
    constexpr int RU = 8;  // rows in flight per warp: their loads are issued together (latency), then consumed
    constexpr int NQ = (FOUT + 31) / 32;
    for (int r0 = rbeg; r0 < rend; r0 += RU) {
      float gl[RU][NQ], xv[RU][NC];
#pragma unroll
      for (int u = 0; u < RU; ++u) {
        const int r = r0 + u;
        const bool live = r < rend;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
          const int o = q * 32 + lane;
          gl[u][q] = (live && o < Fout) ? a.G[(int64_t)r * a.ldg + (int64_t)g * Fout + o] : 0.f;
        }
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int i = c * 32 + lane;
          xv[u][c] = (live && i < Fin) ? a.X[(int64_t)r * a.ldx + (int64_t)g * Fin + i] : 0.f;
        }
      }
#pragma unroll
      for (int u = 0; u < RU; ++u) {
#pragma unroll
        for (int q = 0; q < NQ; ++q) bacc[q] += gl[u][q];
#pragma unroll
        for (int c = 0; c < NC; ++c) xv[u][c] = (xv[u][c] == xv[u][c]) ? xv[u][c] : 0.f;  // NaN input row <=> zero
                                                                // gradient row (Fout rule): drop instead of 0 * NaN
#pragma unroll
        for (int o = 0; o < FOUT; ++o) {
          const float gv = __shfl_sync(0xffffffffu, gl[u][o / 32], o % 32);
#pragma unroll
          for (int c = 0; c < NC; ++c) acc[o][c] = fmaf(gv, xv[u][c], acc[o][c]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < FOUT; ++o) {
      if (o < Fout) {
#pragma unroll
        for (int c = 0; c < NC; ++c) {
          const int i = c * 32 + lane;
          if (i < Fin) {
            const int e = a.w_layout == 0 ? (g * Fout + o) * Fin + i : (g * Fin + i) * Fout + o;
            mine[e] = acc[o][c];
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < (FOUT + 31) / 32; ++q) {
      const int o = q * 32 + lane;
      if (o < Fout) mine[nW + g * Fout + o] = bacc[q];
    }
  }
  __syncthreads();
  float* part = scratch + (size_t)blockIdx.x * E;
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float s = 0.f;
#pragma unroll 8
    for (int w = 0; w < NW; ++w) s += wsm[(size_t)w * E + e];
    part[e] = s;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    float s = 0.f;
#pragma unroll 4
    for (unsigned c = 0; c < gridDim.x; ++c) s += __ldcg(scratch + (size_t)c * E + e);
    if (e < nW) {
      a.dW[e] = a.accumulate ? a.dW[e] + s : s;
    } else if (a.dbias) {
      a.dbias[e - nW] = a.accumulate ? a.dbias[e - nW] + s : s;
    }
  }
  if (threadIdx.x == 0) *ticket = 0u;  // ready for the next launch / graph replay
}

static inline int wgrad_fast_threads(const drgnn_linear_wgrad_args& a) {
  // 0: not eligible.  1024 threads (32 warp strips) when the strips fit shared memory and 64 registers
  // hold the strip; 512 threads for the 64-accumulator shapes
  if (a.Fout > 64 || a.Fin > 64) return 0;
  const int nc = (a.Fin + 31) / 32;
  const int fo = a.Fout <= 16 ? 16 : (a.Fout <= 32 ? 32 : 64);
  if (fo * nc > 64) return 0;
  const int64_t E = (int64_t)a.groups * a.Fout * (a.Fin + 1);
  const int nt = 256;  // measured on B200 (N = 12800, 32 x 32): 256 threads x 64 CTAs 19.6 us; 1024 x 16: 24.6 us
  if (E * (nt / 32) * 4 > 200 * 1024) return 0;
  return nt;
}
static inline int wgrad_fast_ctas(int rows) {
  int c = (rows + 127) / 128;  // >= 16 rows per warp before another CTA is worth its partial
  if (c > WGF_MAX_CTAS) c = WGF_MAX_CTAS;
  if (c < 1) c = 1;
  return c;
}

template <int FOUT, int NC, int NT>
static int launch_wgrad_fast(const drgnn_linear_wgrad_args& a, int ctas, float* scratch, unsigned* ticket, cudaStream_t st) {
  const size_t smem = (size_t)(NT / 32) * a.groups * a.Fout * (a.Fin + 1) * sizeof(float);
  static thread_local size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    DRGNN_CHECK_CUDA(cudaFuncSetAttribute(wgrad_warp_kernel<FOUT, NC, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          204 * 1024));
    configured = 204 * 1024;
  }
  wgrad_warp_kernel<FOUT, NC, NT><<<ctas, NT, smem, st>>>(a, scratch, ticket);
  DRGNN_CHECK_LAUNCH("wgrad_warp_kernel");
  return DRGNN_OK;
}

}  // namespace drgnn

using namespace drgnn;

extern "C" int drgnn_linear(const drgnn_linear_args* a, void* stream) {
  DRGNN_REQUIRE(a != nullptr, "linear: args is NULL");
  DRGNN_REQUIRE(a->rows >= 0 && a->Fin > 0 && a->Fout > 0 && a->groups > 0, "linear: bad sizes (rows=%d Fin=%d Fout=%d groups=%d)",
                a->rows, a->Fin, a->Fout, a->groups);
  DRGNN_REQUIRE(a->X && a->W && a->Y, "linear: NULL pointer");
  DRGNN_REQUIRE(a->w_layout == 0 || a->w_layout == 1, "linear: bad w_layout %d", a->w_layout);
  DRGNN_REQUIRE(a->math >= 0 && a->math <= 2, "linear: bad math mode %d", a->math);
  DRGNN_REQUIRE(a->ldx >= a->groups * a->Fin && a->ldy >= a->groups * a->Fout, "linear: leading dimension too small");
  DRGNN_REQUIRE(!a->out_mask || a->ld_mask >= a->groups * a->Fout, "linear: ld_mask too small");
  if (a->rows == 0) return DRGNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int col_tiles = (a->Fout + LT_C - 1) / LT_C;
  dim3 grid((a->rows + LT_R - 1) / LT_R, a->groups * col_tiles);
  if (a->math == 2 && drgnn_linear_tcgen05_supported(a)) return drgnn_linear_tcgen05(a, stream);
  if (a->math >= 1) {
    linear_tf32x3_kernel<<<grid, 128, 0, st>>>(*a, col_tiles);
    DRGNN_CHECK_LAUNCH("linear_tf32x3_kernel");
  } else {
    const int tile_cols = a->Fout < LT_C ? a->Fout : LT_C;
    const int cgt = (tile_cols + 3) / 4;
    const int threads = 16 * cgt < 128 ? 128 : 16 * cgt;
    linear_fma_kernel<<<grid, threads, 0, st>>>(*a, col_tiles, cgt);
    DRGNN_CHECK_LAUNCH("linear_fma_kernel");
  }
  return DRGNN_OK;
}

extern "C" int64_t drgnn_linear_wgrad_work_floats(int32_t rows, int32_t Fin, int32_t Fout, int32_t groups) {
  if (rows < 0 || Fin <= 0 || Fout <= 0 || groups <= 0) return DRGNN_ERR_INVALID;
  const int64_t E = (int64_t)groups * Fout * ((int64_t)Fin + 1);
  const int64_t general = (int64_t)wgrad_chunks(rows) * E;
  const int64_t fast = (int64_t)WGF_MAX_CTAS * E;
  return (general > fast ? general : fast) + 4;  // + the ticket counter of the fast path (last 4 floats)
}

extern "C" int drgnn_linear_wgrad(const drgnn_linear_wgrad_args* a, void* stream) {
  DRGNN_REQUIRE(a != nullptr, "linear_wgrad: args is NULL");
  DRGNN_REQUIRE(a->rows >= 0 && a->Fin > 0 && a->Fout > 0 && a->groups > 0, "linear_wgrad: bad sizes");
  DRGNN_REQUIRE(a->X && a->G && a->dW && a->work, "linear_wgrad: NULL pointer");
  DRGNN_REQUIRE(a->w_layout == 0 || a->w_layout == 1, "linear_wgrad: bad w_layout %d", a->w_layout);
  DRGNN_REQUIRE(a->ldx >= a->groups * a->Fin && a->ldg >= a->groups * a->Fout, "linear_wgrad: leading dimension too small");
  const int64_t need = drgnn_linear_wgrad_work_floats(a->rows, a->Fin, a->Fout, a->groups);
  DRGNN_REQUIRE(a->work_floats >= need, "linear_wgrad: workspace too small (%lld < %lld floats)", (long long)a->work_floats,
                (long long)need);
  cudaStream_t st = (cudaStream_t)stream;
  if (wgrad_fast_threads(*a) != 0) {
    // the ticket lives in the last 4 floats of the workspace: zero before the first use (the caller
    // allocates the workspace zero-filled), reset by the kernel itself afterwards
    unsigned* ticket = reinterpret_cast<unsigned*>(a->work + a->work_floats - 4);
    const int ctas = wgrad_fast_ctas(a->rows);
    const int nc = (a->Fin + 31) / 32;
    if (a->Fout <= 16) return nc == 1 ? launch_wgrad_fast<16, 1, 256>(*a, ctas, a->work, ticket, st)
                                      : launch_wgrad_fast<16, 2, 256>(*a, ctas, a->work, ticket, st);
    if (a->Fout <= 32) return nc == 1 ? launch_wgrad_fast<32, 1, 256>(*a, ctas, a->work, ticket, st)
                                      : launch_wgrad_fast<32, 2, 256>(*a, ctas, a->work, ticket, st);
    return launch_wgrad_fast<64, 1, 256>(*a, ctas, a->work, ticket, st);
  }
  const int n_chunks = wgrad_chunks(a->rows);
  const int o_tiles = (a->Fout + WG_O - 1) / WG_O, k_tiles = (a->Fin + WG_K - 1) / WG_K;
  dim3 grid(n_chunks, a->groups * o_tiles * k_tiles);
  wgrad_partial_kernel<<<grid, 256, 0, st>>>(*a, n_chunks, o_tiles, k_tiles);
  DRGNN_CHECK_LAUNCH("wgrad_partial_kernel");
  const int64_t total = (int64_t)a->groups * a->Fout * ((int64_t)a->Fin + 1);
  wgrad_reduce_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(*a, n_chunks);
  DRGNN_CHECK_LAUNCH("wgrad_reduce_kernel");
  return DRGNN_OK;
}
