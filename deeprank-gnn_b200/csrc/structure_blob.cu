// Structure pass of the per-graph fused kernels: ONE launch, one CTA per graph, output = the graph's
// structure blob (include/drgnn.h, drgnn_structure_io.blob) and nothing else.
//
// Same results as graph_local_kernel (structure.cu), i.e. bit-exact replacements of
//   get_preloaded_cluster / consecutive_cluster   community_pooling.py:25-30, 197; ginet.py:114,129
//   pool_edge + coalesce (structure only)          community_pooling.py:204-205
//   the dst-sorted CSR the aggregation reads instead of x[col] + scatter_add (ginet.py:57-71)
// but built for one CTA per graph with everything in shared memory: the sorted lists of the COARSENED graph are
// read off bitmaps instead of being produced by counting sorts -
//   pooled row r / pooled column c  = set bits of bm[r][.] / bmT[c][.] over pooled ids   (sorted, unique)
//   members of cluster k / q        = set bits of mem0[k][.] / mem1[q][.] over node ids  (ascending)
// (bitmaps sized by the host's per-graph cluster bounds max_k / max_q, checked in the kernel), and the level-0 CSR
// (rows by destination, ascending edge id inside a row = the CPU scatter order) comes from an atomic slot
// assignment followed by a rank sweep inside every row (an edge's place = the number of smaller edge ids in its
// row: deterministic, no per-row bitmap over the edges - that bitmap was n x m / 32 words and kept graphs beyond
// ~250 nodes out of this pass).  The whole pass is: load -> two presence bitmaps + popcount prefixes (dense relabel of
// both cluster levels) -> ONE scatter sweep -> ONE exclusive scan over all row counts -> emit sweeps.  About a dozen
// CTA barriers in total; the counting-sort pass needs ~60.  Nothing cross-graph is computed: the blob holds
// graph-local indices and the fused kernels need no global offsets (no finalize launch).
#include <limits.h>

#include "common.cuh"

namespace drgnn {

static constexpr int SB_THREADS = 512;
static constexpr int SB_WARPS = SB_THREADS / 32;
static constexpr int SB_CAP_WORDS = 1024;        // presence bitmap: cluster-id range of one graph <= 32768

__device__ unsigned long long g_bphase[16];
#define DRGNN_BPHASE(i)                                                         \
  do {                                                                          \
    if (blockIdx.x == 0 && threadIdx.x == 0) g_bphase[i] = (unsigned long long)clock64(); \
  } while (0)

struct BlobPlan {   // word offsets into dynamic shared memory
  int erow, ecol, id0, id1, dense0, dense1, cb0, cb1, cp0, cp1, bm, bmT, mem0, mem1, cnt, ea, eord, eun, xs, bar, total;
  int xs_words;       // > 0: the graph's feature tile is staged for the first aggregation (one bulk copy)
  int max_k, max_q;   // per-graph bounds of the level-0 / level-1 cluster counts the bitmaps are sized for
  int kwk;   // row stride of bm / bmT / mem1 (bitmaps over pooled-node ids): ceil(max_k/32), made odd
  int kwn;   // row stride of mem0 (bitmap over node ids): ceil(max_n/32), made odd
};
__host__ __device__ inline int sb_up4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline int sb_odd_words(int bits) {   // words for `bits` bits, odd and strictly larger (conflict-free row walks)
  int w = ((bits + 31) >> 5) | 1;
  if (w == ((bits + 31) >> 5)) w += 2;
  return w;
}
__host__ __device__ inline BlobPlan blob_plan(int max_n, int max_e, int max_k, int max_q, int weights, int x_words = 0) {
  BlobPlan p;
  int o = 0;
  auto take = [&](int words) { const int at = o; o += sb_up4(words); return at; };
  if (max_k <= 0 || max_k > max_n) max_k = max_n;
  if (max_q <= 0 || max_q > max_k) max_q = max_k;
  p.max_k = max_k;
  p.max_q = max_q;
  p.kwk = sb_odd_words(max_k);
  p.kwn = sb_odd_words(max_n);
  p.erow = take((max_e + 1) / 2);
  p.ecol = take((max_e + 1) / 2);
  p.id0 = take(max_n);            // raw ids minus their minimum; later the fill counters of the level-0 rows
  p.id1 = take(max_n);
  p.dense0 = take((max_n + 1) / 2);
  p.dense1 = take((max_n + 1) / 2);
  p.cb0 = take(SB_CAP_WORDS);
  p.cb1 = take(SB_CAP_WORDS);
  p.cp0 = take(SB_CAP_WORDS);
  p.cp1 = take(SB_CAP_WORDS);
  p.bm = take(max_k * p.kwk);
  p.bmT = take(max_k * p.kwk);
  p.mem0 = take(max_k * p.kwn);
  p.mem1 = take(max_q * p.kwk);
  p.cnt = take(max_n + 3 * max_k + max_q + 8);
  p.ea = take(weights ? max_e : 0);   // edge attributes of the graph (sGAT weights; 4 KB at 1000 edges)
  p.eord = take((max_e + 1) / 2);     // edge id of every level-0 CSR slot
  p.eun = take((max_e + 1) / 2);      // ... before the rank sweep (atomic slot order)
  p.xs_words = x_words > 0 ? sb_up4(x_words) : 0;
  p.xs = take(p.xs_words);            // feature tile of the graph (first aggregation)
  p.bar = take(p.xs_words ? 4 : 0);   // its mbarrier
  p.total = o;
  return p;
}

// idx32: 0 = int64 ids (reference tensors), 1 = int32, 2 = uint16 (compact feeder records)
__device__ __forceinline__ long long sb_ld_id(const void* base, int64_t i, int idx32) {
  return idx32 == 2 ? (long long)reinterpret_cast<const uint16_t*>(base)[i]
         : idx32  ? (long long)reinterpret_cast<const int32_t*>(base)[i] : reinterpret_cast<const int64_t*>(base)[i];
}
__device__ __forceinline__ uint32_t sb_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// bulk copy of the graph's feature tile into shared memory (issue by one thread) / wait for it (every thread)
__device__ __forceinline__ void sb_stage_x(float* dst, const float* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sb_smem_u32(bar)));
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb_smem_u32(bar)), "r"(bytes) : "memory");
  if (bytes)
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sb_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(sb_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void sb_wait_x(uint64_t* bar) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "SB_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n"
      "@p bra SB_DONE_%=;\n"
      "bra SB_WAIT_%=;\n"
      "SB_DONE_%=:\n"
      "}\n" ::"r"(sb_smem_u32(bar))
      : "memory");
}

// min / max over the CTA of two value pairs at once (level-0 and level-1 cluster ids); results in red[0..3]
__device__ __noinline__ void sb_minmax2(long long mn0, long long mx0, long long mn1, long long mx1, long long* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const long long a = __shfl_xor_sync(0xffffffffu, mn0, o), b = __shfl_xor_sync(0xffffffffu, mx0, o);
    const long long c = __shfl_xor_sync(0xffffffffu, mn1, o), d = __shfl_xor_sync(0xffffffffu, mx1, o);
    mn0 = a < mn0 ? a : mn0; mx0 = b > mx0 ? b : mx0;
    mn1 = c < mn1 ? c : mn1; mx1 = d > mx1 ? d : mx1;
  }
  if (lane == 0) {
    red[4 + 4 * w + 0] = mn0; red[4 + 4 * w + 1] = mx0; red[4 + 4 * w + 2] = mn1; red[4 + 4 * w + 3] = mx1;
  }
  __syncthreads();
  if (w == 0) {
    mn0 = lane < SB_WARPS ? red[4 + 4 * lane + 0] : LLONG_MAX;
    mx0 = lane < SB_WARPS ? red[4 + 4 * lane + 1] : LLONG_MIN;
    mn1 = lane < SB_WARPS ? red[4 + 4 * lane + 2] : LLONG_MAX;
    mx1 = lane < SB_WARPS ? red[4 + 4 * lane + 3] : LLONG_MIN;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const long long a = __shfl_xor_sync(0xffffffffu, mn0, o), b = __shfl_xor_sync(0xffffffffu, mx0, o);
      const long long c = __shfl_xor_sync(0xffffffffu, mn1, o), d = __shfl_xor_sync(0xffffffffu, mx1, o);
      mn0 = a < mn0 ? a : mn0; mx0 = b > mx0 ? b : mx0;
      mn1 = c < mn1 ? c : mn1; mx1 = d > mx1 ? d : mx1;
    }
    if (lane == 0) {
      red[0] = mn0; red[1] = mx0; red[2] = mn1; red[3] = mx1;
    }
  }
  __syncthreads();
}

// exclusive popcount prefix of `W` presence words (cpre) by one warp (W <= 32) or the CTA; total -> *K
__device__ __forceinline__ void sb_prefix_warp(const uint32_t* cb, int* cp, int W, int* Kout) {
  const int lane = threadIdx.x & 31;
  const int c = lane < W ? __popc(cb[lane]) : 0;
  const int incl = warp_scan_incl(c);
  if (lane < W) cp[lane] = incl - c;
  if (lane == 31) *Kout = incl;
}

__global__ void __launch_bounds__(SB_THREADS) graph_blob_kernel(const drgnn_structure_io io, const BlobPlan P) {
  extern __shared__ __align__(16) uint32_t sb[];
  __shared__ long long red[4 + 4 * SB_WARPS];
  __shared__ int wsum[33];
  __shared__ int sK[2];
  const int T = SB_THREADS, t = threadIdx.x, w = t >> 5;
  const int g = blockIdx.x;
  uint16_t* erow = reinterpret_cast<uint16_t*>(sb + P.erow);
  uint16_t* ecol = reinterpret_cast<uint16_t*>(sb + P.ecol);
  int* id0 = reinterpret_cast<int*>(sb + P.id0);
  int* id1 = reinterpret_cast<int*>(sb + P.id1);
  uint16_t* dense0 = reinterpret_cast<uint16_t*>(sb + P.dense0);
  uint16_t* dense1 = reinterpret_cast<uint16_t*>(sb + P.dense1);
  uint32_t* cb0 = sb + P.cb0; uint32_t* cb1 = sb + P.cb1;
  int* cp0 = reinterpret_cast<int*>(sb + P.cp0);
  int* cp1 = reinterpret_cast<int*>(sb + P.cp1);
  uint32_t* bm = sb + P.bm; uint32_t* bmT = sb + P.bmT; uint32_t* mem0 = sb + P.mem0; uint32_t* mem1 = sb + P.mem1;
  int* cnt = reinterpret_cast<int*>(sb + P.cnt);
  int* fill = id0;                      // fill counters of the level-0 rows (the raw ids are dead by then)
  float* eas = reinterpret_cast<float*>(sb + P.ea);
  uint16_t* eord = reinterpret_cast<uint16_t*>(sb + P.eord);
  uint16_t* eun = reinterpret_cast<uint16_t*>(sb + P.eun);
  const int KWK = P.kwk, KWN = P.kwn;
  DRGNN_BPHASE(0);

  const int n0 = io.node_ptr[g], n = io.node_ptr[g + 1] - n0;
  const int e0 = io.edge_ptr[g], m = io.edge_ptr[g + 1] - e0;
  const int c0 = io.c1_ptr[g];
  int c1len = io.c1_ptr[g + 1] - c0;
  int32_t* bl = io.blob + DRGNN_BLOB_OFFSET(g, n0, e0);
  // edge weights of the lists (sGAT, sGAT.py:76): a float array parallel to the blob
  float* wb = (io.wblob && io.edge_attr) ? io.wblob + DRGNN_BLOB_OFFSET(g, n0, e0) : nullptr;
  const BlobLayout BL = blob_layout(n, m);
  if (t < DRGNN_BLOB_HEADER) bl[t] = 0;
  if (t == 0) {   // the graph's extents for the step kernel (io.gstat doubles as its descriptor table)
    int32_t* gs = io.gstat + 8 * g;
    gs[3] = n0; gs[4] = e0; gs[5] = m; gs[6] = 0; gs[7] = n;
  }
  if (n < 0 || m < 0 || n > io.max_n || m > io.max_e) {   // host bounds violated: header stays incomplete
    if (t == 0) atomicOr(io.status, DRGNN_ST_FUSED_BOUNDS);
    return;
  }
  bool bad1 = c1len > io.max_n || c1len < 0;
  if (bad1) c1len = 0;
  // the feature tile of the first aggregation (step 9) starts its way into shared memory now
  const bool stage_x = io.zin1 != nullptr && P.xs_words > 0;
  float* xs = reinterpret_cast<float*>(sb + P.xs);
  uint64_t* xbar = reinterpret_cast<uint64_t*>(sb + P.bar);
  if (stage_x && t == 0) sb_stage_x(xs, io.x + (int64_t)n0 * io.F, (uint32_t)(n * io.F) * 4u, xbar);
  if (stage_x) __syncthreads();          // the barrier is initialised before anybody may wait on it

  // ---- 0. zero every bitmap (the presence words follow once the id ranges are known)
#pragma unroll 1
  for (int i = t; i < P.max_k * KWK; i += T) {
    bm[i] = 0u; bmT[i] = 0u;
  }
#pragma unroll 1
  for (int i = t; i < P.max_k * KWN; i += T) mem0[i] = 0u;
#pragma unroll 1
  for (int i = t; i < P.max_q * KWK; i += T) mem1[i] = 0u;
  // ---- 1. local edge list; cluster ids of both levels (read from global memory once), their extremes
  // (four edges per thread and sweep: the global loads of a sweep are issued together, then checked and stored -
  // one dependent load per loop iteration made this phase 18 k cycles for a graph of 8000 edges)
  if (io.edge16 == 2) {
    // compact feeder batches: m / 2 undirected pairs of uint16 graph-local ids; pair k gives edge k and its mirror
    // m / 2 + k.  An odd edge count cannot be two mirrored halves: every edge is flagged below.
    const uint16_t* ei = reinterpret_cast<const uint16_t*>(io.edge_index);
    const int mh = m >> 1;
    const int64_t h0 = e0 >> 1, Eh = (int64_t)io.E >> 1;
    const bool odd = ((m | e0) & 1) != 0;
#pragma unroll 1
    for (int k0 = t; k0 < mh; k0 += 4 * T) {
      int a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u * T;
        a[u] = k < mh ? (int)ei[h0 + k] : 0;
        b[u] = k < mh ? (int)ei[Eh + h0 + k] : 0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int k = k0 + u * T;
        if (k < mh) {
          int r = a[u], c = b[u];
          if (odd || r >= n || c >= n) {
            atomicOr(io.status, DRGNN_ST_EDGE_OUTSIDE_GRAPH);
            r = 0;
            c = 0;
          }
          erow[k] = (uint16_t)r;      ecol[k] = (uint16_t)c;
          erow[mh + k] = (uint16_t)c; ecol[mh + k] = (uint16_t)r;
        }
      }
    }
    if (odd && t == 0) { erow[m - 1] = 0; ecol[m - 1] = 0; atomicOr(io.status, DRGNN_ST_EDGE_OUTSIDE_GRAPH); }
  } else {
#pragma unroll 1
    for (int e00 = t; e00 < m; e00 += 4 * T) {
      long long rr[4], cc[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e00 + u * T;
        rr[u] = cc[u] = 0;
        if (e < m) {
          if (io.edge16) {   // uint16 graph-local ids, both directions stored
            rr[u] = reinterpret_cast<const uint16_t*>(io.edge_index)[(int64_t)e0 + e];
            cc[u] = reinterpret_cast<const uint16_t*>(io.edge_index)[(int64_t)io.E + e0 + e];
          } else {
            rr[u] = sb_ld_id(io.edge_index, (int64_t)e0 + e, io.idx32) - n0;
            cc[u] = sb_ld_id(io.edge_index, (int64_t)io.E + e0 + e, io.idx32) - n0;
          }
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = e00 + u * T;
        if (e < m) {
          long long r = rr[u], c = cc[u];
          if (r < 0 || r >= n || c < 0 || c >= n) {
            atomicOr(io.status, DRGNN_ST_EDGE_OUTSIDE_GRAPH);
            r = 0;
            c = 0;
          }
          erow[e] = (uint16_t)r;
          ecol[e] = (uint16_t)c;
        }
      }
    }
  }
  if (wb) {   // staged once: the weight sums read shared memory
#pragma unroll 4
    for (int e = t; e < m; e += T) eas[e] = io.edge_attr[(int64_t)(e0 + e) * io.ne];
  }
  long long mn0 = LLONG_MAX, mx0 = LLONG_MIN, mn1 = LLONG_MAX, mx1 = LLONG_MIN;
  // raw ids are kept in registers for graphs of up to 2 * T nodes (else re-read)
  long long r0a = 0, r0b = 0, r1a = 0, r1b = 0;
  if (t < n) { r0a = sb_ld_id(io.cluster0, (int64_t)n0 + t, io.idx32); mn0 = r0a; mx0 = r0a; }
  if (t + T < n) { r0b = sb_ld_id(io.cluster0, (int64_t)n0 + t + T, io.idx32); mn0 = r0b < mn0 ? r0b : mn0; mx0 = r0b > mx0 ? r0b : mx0; }
#pragma unroll 1
  for (int i = t + 2 * T; i < n; i += T) {
    const long long v = sb_ld_id(io.cluster0, (int64_t)n0 + i, io.idx32);
    mn0 = v < mn0 ? v : mn0; mx0 = v > mx0 ? v : mx0;
  }
  if (t < c1len) { r1a = sb_ld_id(io.cluster1, (int64_t)c0 + t, io.idx32); mn1 = r1a; mx1 = r1a; }
  if (t + T < c1len) { r1b = sb_ld_id(io.cluster1, (int64_t)c0 + t + T, io.idx32); mn1 = r1b < mn1 ? r1b : mn1; mx1 = r1b > mx1 ? r1b : mx1; }
#pragma unroll 1
  for (int i = t + 2 * T; i < c1len; i += T) {
    const long long v = sb_ld_id(io.cluster1, (int64_t)c0 + i, io.idx32);
    mn1 = v < mn1 ? v : mn1; mx1 = v > mx1 ? v : mx1;
  }
  if (io.idx32) {   // int32 ids (packed feeder batches): hardware warp reductions, then 16 values per extreme
    const int a0 = __reduce_min_sync(0xffffffffu, (int)(mn0 > INT_MAX ? INT_MAX : mn0));
    const int b0 = __reduce_max_sync(0xffffffffu, (int)(mx0 < INT_MIN ? INT_MIN : mx0));
    const int a1 = __reduce_min_sync(0xffffffffu, (int)(mn1 > INT_MAX ? INT_MAX : mn1));
    const int b1 = __reduce_max_sync(0xffffffffu, (int)(mx1 < INT_MIN ? INT_MIN : mx1));
    int* r32 = reinterpret_cast<int*>(red);
    if ((t & 31) == 0) { r32[4 * w] = a0; r32[4 * w + 1] = b0; r32[4 * w + 2] = a1; r32[4 * w + 3] = b1; }
    __syncthreads();
    int v0 = INT_MAX, v1 = INT_MIN, v2 = INT_MAX, v3 = INT_MIN;
#pragma unroll
    for (int u = 0; u < SB_WARPS; ++u) {
      v0 = min(v0, r32[4 * u]); v1 = max(v1, r32[4 * u + 1]); v2 = min(v2, r32[4 * u + 2]); v3 = max(v3, r32[4 * u + 3]);
    }
    mn0 = v0; mx0 = v1; mn1 = v2; mx1 = v3;
  } else {
    sb_minmax2(mn0, mx0, mn1, mx1, red);
    mn0 = red[0]; mx0 = red[1]; mn1 = red[2]; mx1 = red[3];
  }
  DRGNN_BPHASE(1);
  // ---- 2. presence bitmaps over the id ranges
  const long long range0 = n > 0 ? mx0 - mn0 + 1 : 0, range1 = c1len > 0 ? mx1 - mn1 + 1 : 0;
  bool bad_range = range0 > (long long)SB_CAP_WORDS * 32 || range1 > (long long)SB_CAP_WORDS * 32;
  if ((n > 0 && mn0 < 0) || (c1len > 0 && mn1 < 0)) {
    if (t == 0) atomicOr(io.status, DRGNN_ST_NEGATIVE_ID);
  }
  if (bad_range) {   // same flag as the full pass; the blob stays incomplete (the step kernel refuses it)
    if (t == 0) atomicOr(io.status, DRGNN_ST_CLUSTER_RANGE);
    if (stage_x) sb_wait_x(xbar);        // no bulk copy may be in flight into a CTA that exits
    return;
  }
  const int W0 = (int)((range0 + 31) >> 5), W1c = (int)((range1 + 31) >> 5);
#pragma unroll 1
  for (int i = t; i < W0; i += T) cb0[i] = 0u;
#pragma unroll 1
  for (int i = t; i < W1c; i += T) cb1[i] = 0u;
  __syncthreads();
  if (t < n) { const int v = (int)(r0a - mn0); id0[t] = v; atomicOr(&cb0[v >> 5], 1u << (v & 31)); }
  if (t + T < n) { const int v = (int)(r0b - mn0); id0[t + T] = v; atomicOr(&cb0[v >> 5], 1u << (v & 31)); }
#pragma unroll 1
  for (int i = t + 2 * T; i < n; i += T) {
    const int v = (int)(sb_ld_id(io.cluster0, (int64_t)n0 + i, io.idx32) - mn0);
    id0[i] = v;
    atomicOr(&cb0[v >> 5], 1u << (v & 31));
  }
  if (t < c1len) { const int v = (int)(r1a - mn1); id1[t] = v; atomicOr(&cb1[v >> 5], 1u << (v & 31)); }
  if (t + T < c1len) { const int v = (int)(r1b - mn1); id1[t + T] = v; atomicOr(&cb1[v >> 5], 1u << (v & 31)); }
#pragma unroll 1
  for (int i = t + 2 * T; i < c1len; i += T) {
    const int v = (int)(sb_ld_id(io.cluster1, (int64_t)c0 + i, io.idx32) - mn1);
    id1[i] = v;
    atomicOr(&cb1[v >> 5], 1u << (v & 31));
  }
  __syncthreads();
  // ---- 3. popcount prefixes -> number of clusters of both levels
  if (W0 <= 32 && W1c <= 32) {
    if (w == 0) sb_prefix_warp(cb0, cp0, W0, &sK[0]);
    if (w == 1) sb_prefix_warp(cb1, cp1, W1c, &sK[1]);
    __syncthreads();
  } else {
#pragma unroll 1
    for (int i = t; i < W0; i += T) cp0[i] = __popc(cb0[i]);
#pragma unroll 1
    for (int i = t; i < W1c; i += T) cp1[i] = __popc(cb1[i]);
    __syncthreads();
    const int k0 = block_exclusive_scan(cp0, W0, wsum);
    const int k1 = block_exclusive_scan(cp1, W1c, wsum);
    if (t == 0) { sK[0] = k0; sK[1] = k1; }
    __syncthreads();
  }
  const int K = sK[0], K1 = sK[1];
  if (c1len != K || bad1) {
    if (t == 0) atomicOr(io.status, DRGNN_ST_CLUSTER1_LENGTH);
    bad1 = true;
  }
  if (K > P.max_k || K1 > P.max_q) {   // the host's per-graph cluster bounds (bitmap sizes) are violated: blob stays incomplete
    if (t == 0) atomicOr(io.status, DRGNN_ST_FUSED_BOUNDS);
    if (stage_x) sb_wait_x(xbar);
    return;
  }
  DRGNN_BPHASE(2);
  // ---- 4. dense ids (consecutive_cluster restricted to the graph); into the blob as cl0 / cl1
#pragma unroll 1
  for (int i = t; i < n; i += T) {
    const int v = id0[i];
    const int d = cp0[v >> 5] + __popc(cb0[v >> 5] & ((1u << (v & 31)) - 1u));
    dense0[i] = (uint16_t)d;
    bl[BL.cl0 + i] = d;
  }
#pragma unroll 1
  for (int i = t; i < c1len; i += T) {
    const int v = id1[i];
    const int d = cp1[v >> 5] + __popc(cb1[v >> 5] & ((1u << (v & 31)) - 1u));
    dense1[i] = (uint16_t)d;
    if (i < n) bl[BL.cl1 + i] = d;
  }
  __syncthreads();
  // ---- 5. ONE scatter sweep: every sorted list of the blob becomes a bitmap row, and every row is
  // counted on the way (a pooled edge counts when its bit was not set before: coalesce's unique)
  //   cnt: [0, n) level-0 CSR rows | [n, n+K) pooled rows | [n+K, n+2K) pooled columns
  //        [n+2K, n+3K) members of cluster k | [n+3K, n+3K+K1) members of level-1 cluster q
  const int L = n + 3 * K + K1;
#pragma unroll 1
  for (int i = t; i <= L; i += T) cnt[i] = 0;
  __syncthreads();
#pragma unroll 1
  for (int e = t; e < m; e += T) {
    const int r = erow[e], c = ecol[e];
    atomicAdd(&cnt[r], 1);
    const int pr = dense0[r], pc = dense0[c];
    if (pr != pc) {   // remove_self_loops of pool_edge
      const uint32_t bit = 1u << (pc & 31);
      if (!(atomicOr(&bm[pr * KWK + (pc >> 5)], bit) & bit)) {   // first edge of this pooled pair
        atomicOr(&bmT[pc * KWK + (pr >> 5)], 1u << (pr & 31));
        atomicAdd(&cnt[n + pr], 1);
        atomicAdd(&cnt[n + K + pc], 1);
      }
    }
  }
#pragma unroll 1
  for (int i = t; i < n; i += T) {
    const int k = dense0[i];
    atomicOr(&mem0[k * KWN + (i >> 5)], 1u << (i & 31));
    atomicAdd(&cnt[n + 2 * K + k], 1);
    fill[i] = 0;
  }
#pragma unroll 1
  for (int i = t; i < c1len; i += T) {
    const int q = dense1[i];
    atomicOr(&mem1[q * KWK + (i >> 5)], 1u << (i & 31));
    atomicAdd(&cnt[n + 3 * K + q], 1);
  }
  __syncthreads();
  DRGNN_BPHASE(3);
  // ---- 6. ONE exclusive scan over all row counts
  block_exclusive_scan(cnt, L + 1, wsum);
  const int base1 = cnt[n], baseT = cnt[n + K], baseM0 = cnt[n + 2 * K], baseM1 = cnt[n + 3 * K];
  const int E1 = baseT - base1;
  const int KWk = (K + 31) >> 5;
  DRGNN_BPHASE(4);
  // ---- 7. emit.  Level-0 CSR: every edge takes a slot of its row (atomic: any order), then every slot ranks its
  // edge id among the ids of its row - ascending edge id inside a row = the CPU scatter order, deterministic.
#pragma unroll 1
  for (int e = t; e < m; e += T) {
    const int r = erow[e];
    eun[cnt[r] + atomicAdd(&fill[r], 1)] = (uint16_t)e;
  }
  __syncthreads();
#pragma unroll 1
  for (int p = t; p < m; p += T) {
    const int e = eun[p], r = erow[e];
    const int a = cnt[r], b = cnt[r + 1];           // cnt[n] = m closes the last row
    int rank = 0;
#pragma unroll 4
    for (int q = a; q < b; ++q) rank += (int)eun[q] < e ? 1 : 0;
    const int pos = a + rank;
    eord[pos] = (uint16_t)e;       // edge id of the CSR slot: the weight sums / the first aggregation walk a node's edges in slot order
    bl[BL.col0 + pos] = ecol[e];
    if (wb) wb[BL.col0 + pos] = eas[e];
  }
  DRGNN_BPHASE(9);
#pragma unroll 1
  for (int i = t; i < n; i += T) {          // members of the level-0 clusters, ascending node id
    bl[BL.rp0 + i] = cnt[i];
    const int k = dense0[i], wi = i >> 5;
    const uint32_t* row = mem0 + k * KWN;
    int pos = cnt[n + 2 * K + k] - baseM0 + __popc(row[wi] & ((1u << (i & 31)) - 1u));
#pragma unroll 1
    for (int ww = 0; ww < wi; ++ww) pos += __popc(row[ww]);
    bl[BL.cmem0 + pos] = i;
  }
#pragma unroll 1
  for (int i = t; i < c1len; i += T) {      // members of the level-1 clusters, ascending pooled-node id
    const int q = dense1[i], wi = i >> 5;
    const uint32_t* row = mem1 + q * KWK;
    int pos = cnt[n + 3 * K + q] - baseM1 + __popc(row[wi] & ((1u << (i & 31)) - 1u));
#pragma unroll 1
    for (int ww = 0; ww < wi; ++ww) pos += __popc(row[ww]);
    bl[BL.cmem1 + pos] = i;
  }
#pragma unroll 1
  for (int k = t; k < K; k += T) {          // row pointers of the pooled lists
    bl[BL.rp1 + k] = cnt[n + k] - base1;
    bl[BL.cscp1 + k] = cnt[n + K + k] - baseT;
    bl[BL.cmp0 + k] = cnt[n + 2 * K + k] - baseM0;
  }
#pragma unroll 1
  for (int q = t; q < K1; q += T) bl[BL.cmp1 + q] = cnt[n + 3 * K + q] - baseM1;
  DRGNN_BPHASE(10);
  // pooled rows / columns: a thread per (row, word) emits the set bits of its word (sorted, unique)
#pragma unroll 1
  for (int item = t; item < 2 * K * KWk; item += T) {
    const int which = item >= K * KWk;
    const int it = which ? item - K * KWk : item;
    const int r = it / KWk, wi = it - r * KWk;
    const uint32_t* row = (which ? bmT : bm) + r * KWK;
    int q = which ? cnt[n + K + r] - baseT : cnt[n + r] - base1;
#pragma unroll 1
    for (int ww = 0; ww < wi; ++ww) q += __popc(row[ww]);
    int32_t* dst = bl + (which ? BL.cscr1 : BL.col1);
    uint32_t bits = row[wi];
    while (bits) {
      const int b = __ffs(bits) - 1;
      bits &= bits - 1;
      dst[q++] = wi * 32 + b;
    }
  }
  if (wb) {
    // ---- 8. summed attributes of the merged (coalesced) edges, community_pooling.py:204-205.  A WARP per
    // pooled row, a LANE per pooled edge of the row: all lanes walk the same sequence - the members of the
    // row's cluster (ascending), their level-0 edges (ascending edge id: the fixed order of
    // graph_local_kernel, so both structure passes give bit-identical weights) - with warp-uniform control
    // flow, and each lane adds the attribute when the edge lands in ITS pooled column.  Everything is read
    // from shared memory (the attributes were staged in step 1).
    const int KWn = (n + 31) >> 5;
    const int lane = t & 31;
    __syncthreads();   // eord (the emit sweep above) is complete
    DRGNN_BPHASE(6);
#pragma unroll 1
    for (int r = w; r < K; r += SB_WARPS) {
      const uint32_t* row = bm + r * KWK;
      const uint32_t* mrow = mem0 + r * KWN;
      const int rbase = cnt[n + r] - base1;
      const int rcnt = cnt[n + r + 1] - base1 - rbase;     // cnt[n + K] = baseT = base1 + E1 closes the last row
#pragma unroll 1
      for (int s0 = 0; s0 < rcnt; s0 += 32) {
        int tc = -1;                                        // this lane's pooled column: the (s0 + lane)-th set bit
        {
          int want = s0 + lane;
          if (want < rcnt) {
#pragma unroll 1
            for (int ww = 0; ww < KWk; ++ww) {
              const uint32_t bw = row[ww];
              const int c = __popc(bw);
              if (want < c) {
                tc = ww * 32 + (int)__fns(bw, 0, want + 1);
                break;
              }
              want -= c;
            }
          }
        }
        float acc = 0.f;
#pragma unroll 1
        for (int mw = 0; mw < KWn; ++mw) {
          uint32_t mb = mrow[mw];                           // warp-uniform
          while (mb) {
            const int i = mw * 32 + __ffs(mb) - 1;
            mb &= mb - 1;
            const int a = cnt[i], b = cnt[i + 1];           // CSR slots of node i (cnt[n] = m closes the last row)
#pragma unroll 1
            for (int c0 = a; c0 < b; c0 += 32) {            // the lanes fetch up to 32 edges of the node at once ...
              const int nb_ = min(32, b - c0);
              int pc = -2;
              float val = 0.f;
              if (lane < nb_) {
                const int e = eord[c0 + lane];
                pc = (int)dense0[ecol[e]];
                val = eas[e];
              }
#pragma unroll 1
              for (int j = 0; j < nb_; ++j) {               // ... and add them in ascending edge order
                const int pcj = __shfl_sync(0xffffffffu, pc, j);
                const float vj = __shfl_sync(0xffffffffu, val, j);
                if (pcj == tc) acc += vj;
              }
            }
          }
        }
        if (s0 + lane < rcnt) {
          wb[BL.col1 + rbase + s0 + lane] = acc;
          // the same weight at the edge's pooled-CSC slot: column tc, rank of row r among the rows of bmT[tc]
          const uint32_t* rowT = bmT + tc * KWK;
          int posT = cnt[n + K + tc] - baseT + __popc(rowT[r >> 5] & ((1u << (r & 31)) - 1u));
#pragma unroll 1
          for (int ww = 0; ww < (r >> 5); ++ww) posT += __popc(rowT[ww]);
          wb[BL.cscr1 + posT] = acc;
        }
      }
    }
    DRGNN_BPHASE(7);
  }
  DRGNN_BPHASE(11);
  if (io.zin1) {
    // ---- 9. the first aggregation of the network (optional): the input rows of conv1's dense transform depend on
    // the batch only, not on the weights, so they are computed HERE - on the side stream, while the previous step
    // still computes - and the step kernel stages them ready-made instead of the feature tile:
    //   kind 0 (GINet, ginet.py:57-71)    zin1_i = sum_e x_col
    //   kind 1 (sGAT, sGAT.py:70-92)      zin1_i = [ s_i x_i | (1/max(deg,1)) sum_e a_e x_col | 1 0 0 0 ],  s_i = mean_e a_e
    //   kind 2 (FoutNet, foutnet.py:62-80) zin1_i = [ x_i | (1/deg) sum_e x_col | 1 0 0 0 ]   (deg = 0 -> NaN)
    // Same arithmetic, in the same order (ascending CSR slot = the CPU scatter order), as the aggregation phase of the
    // step kernels (s2_gather / s3_aggregate), so the rows are bit-identical.  The feature rows come from the tile
    // staged in shared memory at the start of the kernel (one bulk copy, in flight during steps 1-8) - from global
    // memory only when the tile does not fit (each row is read deg times: 29 k of 64 k cycles at cfg4).
    __syncthreads();   // eord (the emit sweep above) is complete; cnt holds the CSR row pointers
    const int F = io.F, F4 = F >> 2, ld = io.ld_zin1, kind = io.zin_kind;
    if (stage_x) sb_wait_x(xbar);
    const float* xg = stage_x ? xs : io.x + (int64_t)n0 * F;   // staged tile (shared memory) or the global rows
    float* zg = io.zin1 + (int64_t)n0 * ld;
#pragma unroll 1
    for (int item = t; item < n * F4; item += T) {
      const int i = item / F4, q4 = item - i * F4;
      const int sb_ = cnt[i], se_ = cnt[i + 1];
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      float wsum = 0.f;
      if (kind == 1) {
#pragma unroll 4
        for (int p = sb_; p < se_; ++p) {
          const int e = eord[p];
          const float wv = eas[e];
          const float4 v = *(reinterpret_cast<const float4*>(xg + (int)ecol[e] * F) + q4);
          acc.x = fmaf(wv, v.x, acc.x); acc.y = fmaf(wv, v.y, acc.y); acc.z = fmaf(wv, v.z, acc.z); acc.w = fmaf(wv, v.w, acc.w);
          wsum += wv;
        }
      } else {
#pragma unroll 4
        for (int p = sb_; p < se_; ++p) {
          const float4 v = *(reinterpret_cast<const float4*>(xg + (int)ecol[eord[p]] * F) + q4);
          acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
      }
      float* zr = zg + (int64_t)i * ld;
      if (kind == 0) {
        *reinterpret_cast<float4*>(zr + q4 * 4) = acc;
      } else {
        const int deg = se_ - sb_;
        const float post = kind == 1 ? 1.f / (float)max(deg, 1) : 1.f / (float)deg;
        const float selfc = kind == 1 ? post * wsum : 1.f;
        acc.x *= post; acc.y *= post; acc.z *= post; acc.w *= post;
        const float4 sv = *(reinterpret_cast<const float4*>(xg + i * F) + q4);
        *reinterpret_cast<float4*>(zr + q4 * 4) = make_float4(selfc * sv.x, selfc * sv.y, selfc * sv.z, selfc * sv.w);
        *reinterpret_cast<float4*>(zr + F + q4 * 4) = acc;
        if (q4 == 0) *reinterpret_cast<float4*>(zr + 2 * F) = make_float4(1.f, 0.f, 0.f, 0.f);   // ones column: bias gradient
      }
    }
    DRGNN_BPHASE(8);
  }
  if (t == 0) {   // closing pointers and the header
    bl[BL.rp0 + n] = m;
    bl[BL.rp1 + K] = E1;
    bl[BL.cscp1 + K] = E1;
    bl[BL.cmp0 + K] = n;
    bl[BL.cmp1 + K1] = c1len;
    bl[0] = n; bl[1] = m; bl[2] = K; bl[3] = E1; bl[4] = K1;
    bl[5] = bad1 ? 0 : 1;
    int32_t* gs = io.gstat + 8 * g;
    gs[0] = K; gs[1] = E1; gs[2] = K1;
  }
  DRGNN_BPHASE(5);
}

}  // namespace drgnn

using namespace drgnn;

extern "C" int64_t drgnn_structure_blob_smem_bytes_ex(int32_t max_n, int32_t max_e, int32_t max_k, int32_t max_q,
                                                      int32_t weights, int32_t x_words) {
  if (max_n <= 0 || max_e < 0 || x_words < 0) return DRGNN_ERR_INVALID;
  if (max_n > 16384 || max_e > 65535) return DRGNN_ERR_UNSUPPORTED;   // uint16 ids inside the kernel
  const BlobPlan p = blob_plan(max_n, max_e, max_k, max_q, weights, x_words);
  const int64_t bytes = 4 * (int64_t)p.total;
  if (bytes > device_info().smem_optin - 4096) return DRGNN_ERR_UNSUPPORTED;
  return bytes;
}
extern "C" int64_t drgnn_structure_blob_smem_bytes(int32_t max_n, int32_t max_e) {
  return drgnn_structure_blob_smem_bytes_ex(max_n, max_e, 0, 0, 1, 0);
}

// Blob-only structure pass (the inputs of drgnn_structure_build; outputs: io->blob, io->status,
// io->gstat).  The caller zeroes status.  DRGNN_ERR_UNSUPPORTED when a graph of max_n / max_e does
// not fit the bitmap kernel: run drgnn_structure_build (which also writes the blob) instead.
extern "C" int drgnn_structure_blob(const drgnn_structure_io* io, void* stream) {
  DRGNN_REQUIRE(io != nullptr, "structure_blob: io is NULL");
  DRGNN_REQUIRE(io->B >= 0 && io->N >= 0 && io->E >= 0, "structure_blob: negative size");
  if (io->B == 0) return DRGNN_OK;
  DRGNN_REQUIRE(io->node_ptr && io->edge_ptr && io->edge_index && io->cluster0 && io->c1_ptr && io->cluster1,
                "structure_blob: NULL input (both cluster levels are required)");
  DRGNN_REQUIRE(io->blob && io->status && io->gstat, "structure_blob: NULL output");
  DRGNN_REQUIRE(!io->wblob || !io->edge_attr || io->ne >= 1, "structure_blob: edge weights requested but ne == 0");
  DRGNN_REQUIRE(!io->zin1 || (io->x && io->F > 0 && io->F % 4 == 0 && io->zin_kind >= 0 && io->zin_kind <= 2 &&
                               io->ld_zin1 % 4 == 0 && io->ld_zin1 >= (io->zin_kind ? 2 * io->F + 4 : io->F)),
                "structure_blob: first aggregation requested with an invalid x / F / ld_zin1 / zin_kind");
  DRGNN_REQUIRE(!io->zin1 || io->zin_kind != 1 || (io->wblob && io->edge_attr), "structure_blob: the sGAT aggregation needs edge_attr and wblob");
  const int weights = (io->wblob && io->edge_attr) ? 1 : 0;
  // the feature tile of the first aggregation is staged in shared memory when it fits next to the rest
  int x_words = io->zin1 ? io->max_n * io->F : 0;
  int64_t smem = drgnn_structure_blob_smem_bytes_ex(io->max_n, io->max_e, io->max_k, io->max_q, weights, x_words);
  if (smem < 0 && x_words) {
    x_words = 0;
    smem = drgnn_structure_blob_smem_bytes_ex(io->max_n, io->max_e, io->max_k, io->max_q, weights, 0);
  }
  DRGNN_REQUIRE(!io->zin1 || ((uintptr_t)io->x % 16) == 0, "structure_blob: x must be 16-byte aligned");
  if (smem < 0)
    return fail(DRGNN_ERR_UNSUPPORTED, "structure_blob: a graph with %d nodes / %d edges / %d clusters does not fit the bitmap kernel",
                io->max_n, io->max_e, io->max_k);
  static thread_local int64_t configured = -1;
  if (smem > configured) {
    DRGNN_CHECK_CUDA(cudaFuncSetAttribute(graph_blob_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)device_info().smem_optin - 4096));
    configured = device_info().smem_optin - 4096;
  }
  const BlobPlan plan = blob_plan(io->max_n, io->max_e, io->max_k, io->max_q, weights, x_words);
  graph_blob_kernel<<<io->B, SB_THREADS, smem, (cudaStream_t)stream>>>(*io, plan);
  DRGNN_CHECK_LAUNCH("graph_blob_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_debug_blob_cycles(uint64_t* out16) {
  DRGNN_REQUIRE(out16 != nullptr, "debug_blob_cycles: NULL");
  DRGNN_CHECK_CUDA(cudaMemcpyFromSymbol(out16, g_bphase, sizeof(unsigned long long) * 16));
  return DRGNN_OK;
}
