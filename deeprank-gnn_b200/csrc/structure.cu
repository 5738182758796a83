// Structure pass: all integer work that depends only on the batch, one CTA per graph.
//
// Replaces (bit-exact) for a whole mini-batch:
//   get_preloaded_cluster               deeprank_gnn/community_pooling.py:25-30
//   PyG consecutive_cluster             community_pooling.py:197, max_pool_x (ginet.py:114,129)
//   PyG pool_edge + torch_sparse coalesce   community_pooling.py:204-205
//   PyG pool_batch                      community_pooling.py:222-224
// and builds the dst-sorted CSR / src-sorted CSC the aggregation kernels use instead of
// x[col] gathers + scatter_add (ginet.py:57-71).
//
// Kernel A (graph_local_kernel): everything local to a graph, in shared memory:
//   stable counting sorts (CSR, CSC, members-of-cluster), dense relabel of cluster ids via
//   a presence bitmap + popcount prefix, coarsened edge list via a per-warp column bitmap
//   (emits each pooled row sorted & unique, so no sort / dedupe pass is needed), summed
//   edge attributes in a fixed order (deterministic), level-1 relabel.
// Kernel B (graph_finalize_kernel): cross-graph exclusive offsets (K0, E1, K1) and the
//   compaction of the locally-indexed results into their final global positions.
#include <limits.h>

#include "common.cuh"

namespace drgnn {

// Diagnostic: SM clock at the section boundaries of the CTA of graph 0 (drgnn_debug_structure_cycles).
__device__ unsigned long long g_sphase[32];
#define DRGNN_SPHASE(i)                                                         \
  do {                                                                          \
    if (blockIdx.x == 0 && threadIdx.x == 0) g_sphase[i] = (unsigned long long)clock64(); \
  } while (0)

static constexpr int kThreads = 512;
static constexpr int kWarps = kThreads / 32;
static constexpr int kCapWords = 1024;
static constexpr int kStaticSmemReserve = 2048;  // static __shared__ of the kernels counts against the opt-in limit  // presence bitmap: cluster-id range per graph <= 32768

struct SmemPlan {
  // byte offsets into dynamic shared memory
  int erow, ecol, slotR, slotC, prow, ptrR, ptrC, dense0, mptr, mem, rowptr1, cbits, cpre, bm, hist, idc,
      dense1, mptr1, mem1, total;
  int bm_words;  // capacity of the coarsening bitmap (rows of ceil(K/32) words, processed in rounds)
  int nchunk;    // chunks (= participating warps) of the stable counting sort
};

__host__ __device__ inline int align16(int x) { return (x + 15) & ~15; }

// edge_index / cluster ids arrive as int64 (reference tensors) or int32 (packed feeder batches)
// idx32: 0 = int64 ids (reference tensors), 1 = int32, 2 = uint16 (compact feeder records)
__device__ __forceinline__ long long ld_id(const void* base, int64_t i, int idx32) {
  return idx32 == 2 ? (long long)reinterpret_cast<const uint16_t*>(base)[i]
         : idx32  ? (long long)reinterpret_cast<const int32_t*>(base)[i] : reinterpret_cast<const int64_t*>(base)[i];
}

__host__ __device__ inline SmemPlan make_plan(int max_n, int max_e, int max_c1) {
  SmemPlan p;
  int o = 0;
  const int n1 = max_n + 1, c1 = max_c1 + 1;
  const int w1 = (max_n + 31) / 32;
  long long bmw = (long long)max_n * w1;       // every pooled row at once when it is small ...
  if (bmw > 8192) bmw = 8192;                  // ... else rounds of 32 KB
  if (bmw < w1) bmw = w1;
  p.bm_words = (int)bmw;
  p.nchunk = kWarps;                           // histogram [keys][chunks]: shrink the chunk count for huge graphs
  while (p.nchunk > 1 && (long long)n1 * (p.nchunk + 1) * 4 > 48 * 1024) p.nchunk >>= 1;
  p.erow = o;    o = align16(o + 2 * max_e);
  p.ecol = o;    o = align16(o + 2 * max_e);
  p.slotR = o;   o = align16(o + 2 * max_e);
  p.slotC = o;   o = align16(o + 2 * max_e);   // later reused as pooled col list
  p.prow = o;    o = align16(o + 2 * max_e);
  p.ptrR = o;    o = align16(o + 4 * n1);
  p.ptrC = o;    o = align16(o + 4 * n1);      // later reused for the pooled CSC pointers
  p.dense0 = o;  o = align16(o + 2 * max_n);
  p.mptr = o;    o = align16(o + 4 * n1);
  p.mem = o;     o = align16(o + 2 * max_n);
  p.rowptr1 = o; o = align16(o + 4 * n1);
  p.cbits = o;   o = align16(o + 4 * kCapWords);
  p.cpre = o;    o = align16(o + 4 * kCapWords);
  p.bm = o;      o = align16(o + 4 * p.bm_words);
  p.hist = o;    o = align16(o + 4 * (n1 * (p.nchunk + 1) + 4));
  p.idc = o;     o = align16(o + 8 * max_n);
  p.dense1 = o;  o = align16(o + 2 * max_c1);
  p.mptr1 = o;   o = align16(o + 4 * c1);
  p.mem1 = o;    o = align16(o + 2 * max_c1);
  p.total = o;
  return p;
}

// ---------------------------------------------------------------------------------------
// Stable counting sort: slot[ptr[k] .. ptr[k+1]) = indices e (ascending) with key[e] == k.
// The elements are cut into `nchunk` contiguous chunks, one per warp.  Pass 1 counts the keys of a
// chunk into hist[key][chunk] (shared-memory atomics, almost never colliding), a per-key prefix over
// the chunks plus a scan over the keys give the first slot of every (key, chunk), pass 2 hands out
// the slots of a group in arrival order and a checking sweep repairs the rare inversions (equal keys
// inside one warp instruction).  match.any would give the rank directly but costs ~1000 cycles per
// call on 32 distinct keys (tools/micro/lat.cu).  The result is the unique stable order.
// ---------------------------------------------------------------------------------------
// In-place exclusive scan of a[0..len) for len <= blockDim.x (one element per thread, two barriers);
// longer arrays take the chunked block_exclusive_scan.  Returns the total.  All threads must call.
__device__ __noinline__ int block_scan_small(int* a, int len, int* wsum) {
  const int T = blockDim.x, t = threadIdx.x;
  if (len > T) return block_exclusive_scan(a, len, wsum);
  const int v = t < len ? a[t] : 0;
  const int incl = warp_scan_incl(v);
  if (lane_id() == 31) wsum[warp_id()] = incl;
  __syncthreads();
  if (warp_id() == 0) {
    const int nw = T >> 5;
    const int x = lane_id() < nw ? wsum[lane_id()] : 0;
    const int xi = warp_scan_incl(x);
    __syncwarp();
    wsum[lane_id()] = xi - x;
    if (lane_id() == 31) wsum[32] = xi;
  }
  __syncthreads();
  if (t < len) a[t] = wsum[warp_id()] + incl - v;
  const int total = wsum[32];
  __syncthreads();
  return total;
}

__device__ __noinline__ void csr_build(const uint16_t* key, int m, int n, int* ptr, uint16_t* slot, int* hist, int nchunk, int* wsum) {
  const int T = blockDim.x, t = threadIdx.x, lane = lane_id(), w = warp_id();
  const int hs = nchunk + 1;             // row stride of hist[key][chunk]: odd => conflict-free both ways
  {
    int4* h4 = reinterpret_cast<int4*>(hist);
    const int H4 = (n * hs + 3) >> 2;
#pragma unroll 1
    for (int i = t; i < H4; i += T) h4[i] = make_int4(0, 0, 0, 0);
  }
  __syncthreads();
  int chunk = (m + nchunk - 1) / nchunk;
  chunk = (chunk + 31) & ~31;
  const int beg = min(m, w * chunk), end = (w < nchunk) ? min(m, beg + chunk) : beg;
#pragma unroll 1
  for (int e = beg + lane; e < end; e += 32) atomicAdd(&hist[(int)key[e] * hs + w], 1);
  __syncthreads();
  // per key (one thread): exclusive prefix of its chunk counts in place, total into ptr[key]
#pragma unroll 1
  for (int k = t; k < n; k += T) {
    int* hk = hist + k * hs;
    int run = 0;
#pragma unroll 4
    for (int c = 0; c < nchunk; ++c) {
      const int v = hk[c];
      hk[c] = run;
      run += v;
    }
    ptr[k] = run;
  }
  if (t == 0) ptr[n] = 0;
  __syncthreads();
  block_scan_small(ptr, n + 1, wsum);    // ptr[k] = first slot of key k, ptr[n] = m
#pragma unroll 1
  for (int e = beg + lane; e < end; e += 32) {
    const int k = key[e];
    slot[ptr[k] + atomicAdd(&hist[k * hs + w], 1)] = (uint16_t)e;
  }
  __syncthreads();
  // Elements of one (key, chunk) group that met in the same warp instruction got their slots in the
  // order the hardware resolved the colliding atomics: a segmented odd-even transposition restores
  // ascending e (usually nothing to swap: one checking sweep).
  for (;;) {
    int swapped = 0;
#pragma unroll 1
    for (int parity = 0; parity < 2; ++parity) {
#pragma unroll 1
      for (int p = 2 * t + parity; p + 1 < m; p += 2 * T) {
        const uint16_t a = slot[p], b = slot[p + 1];
        if (a > b && key[a] == key[b]) {
          slot[p] = b;
          slot[p + 1] = a;
          swapped = 1;
        }
      }
      __syncthreads();
    }
    if (!__syncthreads_or(swapped)) break;
  }
}

// ---------------------------------------------------------------------------------------
// Dense relabel of `n` int64 ids (sorted-unique rank, i.e. consecutive_cluster's inverse
// restricted to one graph).  Returns K; writes min / max of the raw ids.
// ---------------------------------------------------------------------------------------
__device__ __noinline__ int relabel(const void* ids_base, int64_t ids_off, int idx32, int n, uint16_t* dense, uint32_t* cbits, int* cpre, int* wsum,
                       long long* red, long long* idc, int32_t* status, long long* out_min, long long* out_max) {
  const int T = blockDim.x, t = threadIdx.x, lane = lane_id(), w = warp_id(), NWARP = T >> 5;
  long long mn = LLONG_MAX, mx = LLONG_MIN;
#pragma unroll 1
  for (int i = t; i < n; i += T) {   // the ids are read from global memory once; idc[i] is re-read by the same thread only
    long long v = ld_id(ids_base, ids_off + i, idx32);
    idc[i] = v;
    mn = v < mn ? v : mn;
    mx = v > mx ? v : mx;
  }
  if (idx32) {   // int32 ids (packed feeder batches): hardware warp reductions
    const int a = __reduce_min_sync(0xffffffffu, (int)(mn > INT_MAX ? INT_MAX : mn));
    const int b = __reduce_max_sync(0xffffffffu, (int)(mx < INT_MIN ? INT_MIN : mx));
    mn = a;
    mx = b;
  } else {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
      mn = a < mn ? a : mn;
      mx = b > mx ? b : mx;
    }
  }
  if (lane == 0) {
    red[2 * w] = mn;
    red[2 * w + 1] = mx;
  }
  __syncthreads();
  if (w == 0) {   // warp 0 folds the per-warp extremes, result in red[0], red[1]
    mn = lane < NWARP ? red[2 * lane] : LLONG_MAX;
    mx = lane < NWARP ? red[2 * lane + 1] : LLONG_MIN;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      long long a = __shfl_xor_sync(0xffffffffu, mn, o), b = __shfl_xor_sync(0xffffffffu, mx, o);
      mn = a < mn ? a : mn;
      mx = b > mx ? b : mx;
    }
    __syncwarp();
    if (lane == 0) {
      red[0] = mn;
      red[1] = mx;
    }
  }
  __syncthreads();
  mn = red[0];
  mx = red[1];
  if (n == 0) { mn = LLONG_MAX; mx = LLONG_MIN; }
  *out_min = mn;
  *out_max = mx;
  if (n == 0) {
    __syncthreads();
    return 0;
  }
  long long range = mx - mn + 1;
  if (mn < 0 && t == 0) atomicOr(status, DRGNN_ST_NEGATIVE_ID);
  if (range > (long long)kCapWords * 32) {
    if (t == 0) atomicOr(status, DRGNN_ST_CLUSTER_RANGE);
#pragma unroll 1
    for (int i = t; i < n; i += T) dense[i] = 0;
    __syncthreads();
    return 1;
  }
  const int W = (int)((range + 31) >> 5);
#pragma unroll 1
  for (int i = t; i < W; i += T) cbits[i] = 0u;
  __syncthreads();
#pragma unroll 1
  for (int i = t; i < n; i += T) {
    int v = (int)(idc[i] - mn);
    atomicOr(&cbits[v >> 5], 1u << (v & 31));
  }
  __syncthreads();
  int K;
  if (W <= 32) {   // the usual case (cluster ids of one graph span a few words): one warp, one barrier
    if (w == 0) {
      const int c = lane < W ? __popc(cbits[lane]) : 0;
      const int incl = warp_scan_incl(c);
      if (lane < W) cpre[lane] = incl - c;
      if (lane == 31) wsum[32] = incl;
    }
    __syncthreads();
    K = wsum[32];
  } else {
#pragma unroll 1
    for (int i = t; i < W; i += T) cpre[i] = __popc(cbits[i]);
    __syncthreads();
    K = block_exclusive_scan(cpre, W, wsum);
  }
#pragma unroll 1
  for (int i = t; i < n; i += T) {
    int v = (int)(idc[i] - mn);
    dense[i] = (uint16_t)(cpre[v >> 5] + __popc(cbits[v >> 5] & ((1u << (v & 31)) - 1u)));
  }
  __syncthreads();
  return K;
}

// scratch layout helpers (must match host side)
struct Scratch {
  int32_t *mptr0, *rowptr1, *cscptr1, *mptr1, *mem1;  // node-indexed
  int32_t *col1, *row1, *cscrow1, *csceid1;           // edge-indexed
};
__host__ __device__ inline Scratch make_scratch(const drgnn_structure_io& io) {
  Scratch s;
  const int64_t nb = (int64_t)io.N + io.B + 1;
  s.mptr0 = io.scratch_n;
  s.rowptr1 = io.scratch_n + nb;
  s.cscptr1 = io.scratch_n + 2 * nb;
  s.mptr1 = io.scratch_n + 3 * nb;
  s.mem1 = io.scratch_n + 4 * nb;
  s.col1 = io.scratch_e;
  s.row1 = io.scratch_e + io.E;
  s.cscrow1 = io.scratch_e + 2 * (int64_t)io.E;
  s.csceid1 = io.scratch_e + 3 * (int64_t)io.E;
  return s;
}

__global__ void __launch_bounds__(kThreads) graph_local_kernel(const drgnn_structure_io io) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ int wsum[33];
  __shared__ long long red[2 * kWarps];
  const int T = blockDim.x, t = threadIdx.x;
  const int g = blockIdx.x;
  const SmemPlan P = make_plan(io.max_n, io.max_e, io.L1 > 0 ? io.max_n : 0);
  uint16_t* erow = (uint16_t*)(smem + P.erow);
  uint16_t* ecol = (uint16_t*)(smem + P.ecol);
  uint16_t* slotR = (uint16_t*)(smem + P.slotR);
  uint16_t* slotC = (uint16_t*)(smem + P.slotC);
  uint16_t* prow = (uint16_t*)(smem + P.prow);
  int* ptrR = (int*)(smem + P.ptrR);
  int* ptrC = (int*)(smem + P.ptrC);
  uint16_t* dense0 = (uint16_t*)(smem + P.dense0);
  int* mptr = (int*)(smem + P.mptr);
  uint16_t* mem = (uint16_t*)(smem + P.mem);
  int* rowptr1 = (int*)(smem + P.rowptr1);
  uint32_t* cbits = (uint32_t*)(smem + P.cbits);
  int* cpre = (int*)(smem + P.cpre);
  uint32_t* bm = (uint32_t*)(smem + P.bm);
  int* hist = (int*)(smem + P.hist);
  long long* idc = (long long*)(smem + P.idc);
  const int nchunk = P.nchunk;
  uint16_t* dense1 = (uint16_t*)(smem + P.dense1);
  int* mptr1 = (int*)(smem + P.mptr1);
  uint16_t* mem1 = (uint16_t*)(smem + P.mem1);
  const Scratch S = make_scratch(io);

  const int n0 = io.node_ptr[g], n = io.node_ptr[g + 1] - n0;
  const int e0 = io.edge_ptr[g], m = io.edge_ptr[g + 1] - e0;
  const int ne = io.ne;
  // per-graph structure blob (graph-local indices, one contiguous block; see drgnn.h)
  int32_t* bl = io.blob ? io.blob + DRGNN_BLOB_OFFSET(g, n0, e0) : nullptr;
  // edge weights of the blob's lists (sGAT): a float array parallel to the blob
  float* wb = (io.blob && io.wblob && io.edge_attr) ? io.wblob + DRGNN_BLOB_OFFSET(g, n0, e0) : nullptr;
  const BlobLayout BL = blob_layout(n, m);
  if (bl && t < DRGNN_BLOB_HEADER) bl[t] = 0;   // header[5] (complete) is set at the very end

  DRGNN_SPHASE(0);
  // ---- 1. local edge list ----
#pragma unroll 1
  for (int e = t; e < m; e += T) {
    long long r, c;
    if (io.edge16 == 2) {   // compact feeder batches: m / 2 undirected pairs of uint16 graph-local ids (the second half
      // of the graph's directed edges mirrors the first, DataSet.py:266-269); pairs of the graph start at e0 / 2
      const uint16_t* ei = reinterpret_cast<const uint16_t*>(io.edge_index);
      const int mh = m >> 1, eh = e < mh ? e : e - mh;
      const int64_t at = (int64_t)(e0 >> 1) + eh;
      const long long a = ei[at], b = ei[((int64_t)io.E >> 1) + at];
      r = ((m | e0) & 1) ? -1 : (e < mh ? a : b);   // an odd edge count cannot be two mirrored halves: flagged below
      c = e < mh ? b : a;
    } else if (io.edge16) {   // uint16 graph-local ids, both directions stored
      r = reinterpret_cast<const uint16_t*>(io.edge_index)[(int64_t)e0 + e];
      c = reinterpret_cast<const uint16_t*>(io.edge_index)[(int64_t)io.E + e0 + e];
    } else {
      r = ld_id(io.edge_index, (int64_t)e0 + e, io.idx32) - n0;
      c = ld_id(io.edge_index, (int64_t)io.E + e0 + e, io.idx32) - n0;
    }
    if (r < 0 || r >= n || c < 0 || c >= n) {
      atomicOr(io.status, DRGNN_ST_EDGE_OUTSIDE_GRAPH);
      r = 0;
      c = 0;
    }
    erow[e] = (uint16_t)r;
    ecol[e] = (uint16_t)c;
  }
  __syncthreads();

  DRGNN_SPHASE(1);
  // ---- 2. CSR by destination (row) ----
  csr_build(erow, m, n, ptrR, slotR, hist, nchunk, wsum);
#pragma unroll 1
  for (int i = t; i <= n; i += T) io.rowptr0[n0 + i] = e0 + ptrR[i];
#pragma unroll 1
  for (int p = t; p < m; p += T) {
    int e = slotR[p];
    io.col0[e0 + p] = n0 + ecol[e];
    io.eid0[e0 + p] = e0 + e;
    if (io.w0csr) io.w0csr[e0 + p] = io.edge_attr[(int64_t)(e0 + e) * ne];
    if (bl) bl[BL.col0 + p] = ecol[e];
    if (wb) wb[BL.col0 + p] = io.edge_attr[(int64_t)(e0 + e) * ne];
  }
  if (bl) {
#pragma unroll 1
    for (int i = t; i <= n; i += T) bl[BL.rp0 + i] = ptrR[i];
  }
  DRGNN_SPHASE(2);
  // ---- 3. CSC (transposed graph) ----
  csr_build(ecol, m, n, ptrC, slotC, hist, nchunk, wsum);
#pragma unroll 1
  for (int i = t; i <= n; i += T) io.cscptr0[n0 + i] = e0 + ptrC[i];
#pragma unroll 1
  for (int p = t; p < m; p += T) {
    int e = slotC[p];
    io.cscrow0[e0 + p] = n0 + erow[e];
    io.csceid0[e0 + p] = e0 + e;
    if (io.w0csc) io.w0csc[e0 + p] = io.edge_attr[(int64_t)(e0 + e) * ne];
  }
  __syncthreads();

  DRGNN_SPHASE(3);
  // ---- 4. relabel level-0 clusters ----
  long long cmin, cmax;
  const int K = relabel(io.cluster0, n0, io.idx32, n, dense0, cbits, cpre, wsum, red, idc, io.status, &cmin, &cmax);
#pragma unroll 1
  for (int i = t; i < n; i += T) io.cl0[n0 + i] = dense0[i];  // local; finalize adds the graph offset

  DRGNN_SPHASE(4);
  // ---- 5. members of every cluster (ascending node id) ----
  csr_build(dense0, n, K, mptr, mem, hist, nchunk, wsum);
#pragma unroll 1
  for (int k = t; k <= K; k += T) S.mptr0[n0 + g + k] = mptr[k];
#pragma unroll 1
  for (int p = t; p < n; p += T) io.cmem0[n0 + p] = n0 + mem[p];
  if (bl) {
#pragma unroll 1
    for (int k = t; k <= K; k += T) bl[BL.cmp0 + k] = mptr[k];
#pragma unroll 1
    for (int p = t; p < n; p += T) {
      bl[BL.cmem0 + p] = mem[p];
      bl[BL.cl0 + p] = dense0[p];
    }
  }

  DRGNN_SPHASE(5);
  // ---- 6. coarsened edges: per pooled row a column bitmap -> sorted unique columns ----
  uint16_t* pcol = slotC;  // CSC slots are dead now
  // Bitmap of pooled columns per pooled row, filled by ALL edges in parallel (one atomicOr each);
  // rows are emitted sorted & unique straight from the bits.  When K rows x ceil(K/32) words exceed
  // the bitmap capacity the rows are processed in rounds (count sweep, scan, emit sweep).
  const int W1 = (K + 31) >> 5;
  const int RB = W1 > 0 ? max(1, P.bm_words / W1) : 1;
  const int nround = K > 0 ? (K + RB - 1) / RB : 0;
  for (int pass = 0; pass < 2; ++pass) {
    for (int rd = 0; rd < nround; ++rd) {
      const int r0 = rd * RB, rows = min(RB, K - r0);
      if (pass == 0 || nround > 1) {
#pragma unroll 1
        for (int w = t; w < rows * W1; w += T) bm[w] = 0u;
        __syncthreads();
#pragma unroll 1
        for (int e = t; e < m; e += T) {
          const int pr = dense0[erow[e]], pc = dense0[ecol[e]];
          if (pr != pc && pr >= r0 && pr < r0 + rows)   // remove_self_loops
            atomicOr(&bm[(pr - r0) * W1 + (pc >> 5)], 1u << (pc & 31));
        }
        __syncthreads();
      }
      if (pass == 0) {
#pragma unroll 1
        for (int r = t; r < rows; r += T) {
          int cnt = 0;
          for (int w = 0; w < W1; ++w) cnt += __popc(bm[r * W1 + w]);
          rowptr1[r0 + r] = cnt;
        }
      } else {
#pragma unroll 1
        for (int r = t; r < rows; r += T) {
          int q = rowptr1[r0 + r];
          for (int w = 0; w < W1; ++w) {
            uint32_t bits = bm[r * W1 + w];
            while (bits) {
              const int b = __ffs(bits) - 1;
              bits &= bits - 1;
              pcol[q] = (uint16_t)(w * 32 + b);
              prow[q] = (uint16_t)(r0 + r);
              ++q;
            }
          }
        }
      }
      __syncthreads();
    }
    if (pass == 0) {
      if (t == 0) rowptr1[K] = 0;
      __syncthreads();
      block_exclusive_scan(rowptr1, K + 1, wsum);
    }
  }
  // summed attributes of merged edges: fixed order (members ascending, then edge id); warp per pooled row
  if (io.edge_attr1 != nullptr) {
    const int lane = lane_id();
#pragma unroll 1
    for (int r = warp_id(); r < K; r += kWarps) {
      const int base = rowptr1[r];
      const int cnt = rowptr1[r + 1] - base;
      for (int f = 0; f < ne; ++f) {
#pragma unroll 1
        for (int sidx = lane; sidx < cnt; sidx += 32) {
          const int tc = pcol[base + sidx];
          float acc = 0.f;
          for (int mi = mptr[r]; mi < mptr[r + 1]; ++mi) {
            int i = mem[mi];
            for (int p = ptrR[i]; p < ptrR[i + 1]; ++p) {
              int e = slotR[p];
              if (dense0[ecol[e]] == tc) acc += io.edge_attr[(int64_t)(e0 + e) * ne + f];
            }
          }
          io.scratch_f[(int64_t)(e0 + base + sidx) * ne + f] = acc;
          if (wb && f == 0) wb[BL.col1 + base + sidx] = acc;
        }
      }
    }
  }
  const int E1 = rowptr1[K];
#pragma unroll 1
  for (int r = t; r <= K; r += T) S.rowptr1[n0 + g + r] = rowptr1[r];
#pragma unroll 1
  for (int p = t; p < E1; p += T) {
    S.col1[e0 + p] = pcol[p];
    S.row1[e0 + p] = prow[p];
    if (bl) bl[BL.col1 + p] = pcol[p];
  }
  if (bl) {
#pragma unroll 1
    for (int r = t; r <= K; r += T) bl[BL.rp1 + r] = rowptr1[r];
  }
  __syncthreads();

  DRGNN_SPHASE(6);
  // ---- 7. CSC of the coarsened graph ----
  int* ptrC1 = ptrC;
  uint16_t* slotC1 = slotR;  // level-0 CSR slots are dead now
  csr_build(pcol, E1, K, ptrC1, slotC1, hist, nchunk, wsum);
#pragma unroll 1
  for (int r = t; r <= K; r += T) S.cscptr1[n0 + g + r] = ptrC1[r];
#pragma unroll 1
  for (int p = t; p < E1; p += T) {
    int q = slotC1[p];
    S.cscrow1[e0 + p] = prow[q];
    S.csceid1[e0 + p] = q;
    if (bl) bl[BL.cscr1 + p] = prow[q];
    if (wb) wb[BL.cscr1 + p] = io.scratch_f[(int64_t)(e0 + q) * ne];   // written by this CTA before the barrier above
  }
  if (bl) {
#pragma unroll 1
    for (int r = t; r <= K; r += T) bl[BL.cscp1 + r] = ptrC1[r];
  }
  __syncthreads();

  DRGNN_SPHASE(7);
  // ---- 8. level-1 clustering ----
  int K1 = 0, c1len = 0;
  if (io.cluster1 != nullptr) {
    const int c0 = io.c1_ptr[g];
    c1len = io.c1_ptr[g + 1] - c0;
    if (c1len != K) {
      if (t == 0) atomicOr(io.status, DRGNN_ST_CLUSTER1_LENGTH);
    }
    c1len = min(c1len, io.max_n);  // shared-memory bound (only reachable for invalid input)
    long long mn1, mx1;
    K1 = relabel(io.cluster1, c0, io.idx32, c1len, dense1, cbits, cpre, wsum, red, idc, io.status, &mn1, &mx1);
#pragma unroll 1
    for (int k = t; k < c1len; k += T) io.cl1[c0 + k] = dense1[k];
    csr_build(dense1, c1len, K1, mptr1, mem1, hist, nchunk, wsum);
#pragma unroll 1
    for (int q = t; q <= K1; q += T) S.mptr1[c0 + g + q] = mptr1[q];
#pragma unroll 1
    for (int p = t; p < c1len; p += T) S.mem1[c0 + p] = mem1[p];
    if (bl) {
#pragma unroll 1
      for (int q = t; q <= K1; q += T) bl[BL.cmp1 + q] = mptr1[q];
#pragma unroll 1
      for (int p = t; p < min(c1len, n); p += T) {
        bl[BL.cmem1 + p] = mem1[p];
        bl[BL.cl1 + p] = dense1[p];
      }
    }
  }

  DRGNN_SPHASE(8);
  if (t == 0) {
    int32_t* gs = io.gstat + 8 * g;
    gs[0] = K;
    gs[1] = E1;
    gs[2] = K1;
    gs[3] = (int32_t)(cmin & 0xffffffffll);
    gs[4] = (int32_t)(cmin >> 32);
    gs[5] = (int32_t)(cmax & 0xffffffffll);
    gs[6] = (int32_t)(cmax >> 32);
    gs[7] = n;
    if (bl) {
      bl[0] = n; bl[1] = m; bl[2] = K; bl[3] = E1; bl[4] = K1;
      bl[5] = (io.cluster1 != nullptr && c1len == K) ? 1 : 0;
    }
  }
}

__global__ void __launch_bounds__(256) graph_finalize_kernel(const drgnn_structure_io io) {
  __shared__ int red[3][8];
  __shared__ int offs[3];
  const int T = blockDim.x, t = threadIdx.x;
  const int g = blockIdx.x;
  const Scratch S = make_scratch(io);
  // exclusive offsets over the preceding graphs
  int a = 0, b = 0, c = 0;
#pragma unroll 1
  for (int h = t; h < g; h += T) {
    a += io.gstat[8 * h];
    b += io.gstat[8 * h + 1];
    c += io.gstat[8 * h + 2];
  }
  a = warp_sum(a);
  b = warp_sum(b);
  c = warp_sum(c);
  if (lane_id() == 0) {
    red[0][warp_id()] = a;
    red[1][warp_id()] = b;
    red[2][warp_id()] = c;
  }
  __syncthreads();
  if (t < 3) {
    int s = 0;
    for (int w = 0; w < (T >> 5); ++w) s += red[t][w];
    offs[t] = s;
  }
  __syncthreads();
  const int Koff = offs[0], E1off = offs[1], K1off = offs[2];
  const int32_t* gs = io.gstat + 8 * g;
  const int K = gs[0], E1 = gs[1], K1 = gs[2];
  const int n0 = io.node_ptr[g], n = io.node_ptr[g + 1] - n0;
  const int e0 = io.edge_ptr[g];
  const int ne = io.ne;

  if (!io.clusters_are_local && g > 0 && n > 0 && t == 0) {
    // global ids must increase with the graph id, else sorted-unique order != graph-major order
    int h = g - 1;
    while (h >= 0 && io.gstat[8 * h + 7] == 0) --h;
    if (h >= 0) {
      long long pmax = ((long long)io.gstat[8 * h + 6] << 32) | (uint32_t)io.gstat[8 * h + 5];
      long long cmin = ((long long)gs[4] << 32) | (uint32_t)gs[3];
      if (cmin <= pmax) atomicOr(io.status, DRGNN_ST_CLUSTER_ORDER);
    }
  }

#pragma unroll 1
  for (int i = t; i < n; i += T) {
    int v = io.cl0[n0 + i] + Koff;
    io.cl0[n0 + i] = v;
    if (io.cl0_i64) io.cl0_i64[n0 + i] = v;
  }
#pragma unroll 1
  for (int k = t; k <= K; k += T) {
    io.cmptr0[Koff + k] = n0 + S.mptr0[n0 + g + k];
    io.rowptr1[Koff + k] = E1off + S.rowptr1[n0 + g + k];
    io.cscptr1[Koff + k] = E1off + S.cscptr1[n0 + g + k];
  }
#pragma unroll 1
  for (int k = t; k < K; k += T) {
    io.batch1[Koff + k] = g;
    if (io.batch1_i64) io.batch1_i64[Koff + k] = g;
  }
#pragma unroll 1
  for (int p = t; p < E1; p += T) {
    const int row = Koff + S.row1[e0 + p], col = Koff + S.col1[e0 + p];
    io.col1[E1off + p] = col;
    if (io.edge_index1) {
      io.edge_index1[E1off + p] = row;
      io.edge_index1[(int64_t)io.E + E1off + p] = col;
    }
    if (io.edge_attr1)
      for (int f = 0; f < ne; ++f)
        io.edge_attr1[(int64_t)(E1off + p) * ne + f] = io.scratch_f[(int64_t)(e0 + p) * ne + f];
    const int q = S.csceid1[e0 + p];
    io.cscrow1[E1off + p] = Koff + S.cscrow1[e0 + p];
    io.csceid1[E1off + p] = E1off + q;
    if (io.w1csc) io.w1csc[E1off + p] = io.scratch_f[(int64_t)(e0 + q) * ne];
  }
  if (t == 0) io.kptr0[g] = Koff;

  if (io.cluster1 != nullptr) {
    const int c0 = io.c1_ptr[g], c1len = io.c1_ptr[g + 1] - c0;
#pragma unroll 1
    for (int k = t; k < c1len; k += T) io.cl1[c0 + k] += K1off;
#pragma unroll 1
    for (int q = t; q <= K1; q += T) io.cmptr1[K1off + q] = Koff + S.mptr1[c0 + g + q];
#pragma unroll 1
    for (int p = t; p < c1len; p += T) io.cmem1[Koff + p] = Koff + S.mem1[c0 + p];
#pragma unroll 1
    for (int q = t; q < K1; q += T) {
      io.batch2[K1off + q] = g;
      if (io.batch2_i64) io.batch2_i64[K1off + q] = g;
    }
    if (t == 0) io.kptr1[g] = K1off;
  }
  if (g == io.B - 1 && t == 0) {
    io.kptr0[io.B] = Koff + K;
    if (io.cluster1 != nullptr) io.kptr1[io.B] = K1off + K1;
    io.counts[0] = Koff + K;
    io.counts[1] = E1off + E1;
    io.counts[2] = K1off + K1;
    io.counts[3] = 0;
  }
}

// ---------------------------------------------------------------------------------------
// get_preloaded_cluster stand-alone (community_pooling.py:25-30)
// ---------------------------------------------------------------------------------------
__global__ void seg_max_kernel(const int64_t* cluster, const int32_t* seg_ptr, int B, int64_t* segmax1) {
  __shared__ long long red[8];
  const int g = blockIdx.x;
  long long mx = LLONG_MIN;
  for (int i = seg_ptr[g] + threadIdx.x; i < seg_ptr[g + 1]; i += blockDim.x) {
    long long v = cluster[i];
    mx = v > mx ? v : mx;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    long long a = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = a > mx ? a : mx;
  }
  if (lane_id() == 0) red[warp_id()] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w) mx = red[w] > mx ? red[w] : mx;
    // torch.max of an empty selection raises in the reference; an empty graph contributes 0 here
    segmax1[g] = (seg_ptr[g + 1] > seg_ptr[g]) ? mx + 1 : 0;
  }
}
__global__ void seg_offset_add_kernel(int64_t* cluster, const int32_t* seg_ptr, int B, const int64_t* segmax1) {
  __shared__ long long red[8];
  __shared__ long long off_s;
  const int g = blockIdx.x;
  long long s = 0;
  for (int h = threadIdx.x; h < g; h += blockDim.x) s += segmax1[h];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane_id() == 0) red[warp_id()] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long tot = 0;
    for (int w = 0; w < (blockDim.x >> 5); ++w) tot += red[w];
    off_s = tot;
  }
  __syncthreads();
  const long long off = off_s;
  for (int i = seg_ptr[g] + threadIdx.x; i < seg_ptr[g + 1]; i += blockDim.x) cluster[i] += off;
}

__global__ void ptr_from_sorted_ids_kernel(const int64_t* ids, int n, int B, int32_t* ptr, int32_t* status) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  // boundaries: ptr[g] = first index whose id >= g
  long long prev = (i == 0) ? -1 : ids[i - 1];
  long long cur = (i == n) ? B : ids[i];
  if (i < n && (cur < 0 || cur >= B)) {
    atomicOr(status, 1);
    return;
  }
  if (cur < prev) {
    atomicOr(status, 1);
    return;
  }
  for (long long gidx = prev + 1; gidx <= cur; ++gidx) ptr[gidx] = i;
}

}  // namespace drgnn

using namespace drgnn;

extern "C" int64_t drgnn_structure_smem_bytes(int32_t max_n, int32_t max_e, int32_t max_c1) {
  if (max_n < 0 || max_e < 0) return DRGNN_ERR_INVALID;
  if (max_n > 65535 || max_e > 65535) return DRGNN_ERR_UNSUPPORTED;
  SmemPlan p = make_plan(max_n, max_e, max_c1);
  if (p.total > device_info().smem_optin - kStaticSmemReserve) return DRGNN_ERR_UNSUPPORTED;
  return p.total;
}

extern "C" int drgnn_structure_build(const drgnn_structure_io* io, void* stream) {
  DRGNN_REQUIRE(io != nullptr, "structure_build: io is NULL");
  DRGNN_REQUIRE(io->B >= 0 && io->N >= 0 && io->E >= 0, "structure_build: negative size");
  if (io->B == 0) return DRGNN_OK;
  DRGNN_REQUIRE(io->node_ptr && io->edge_ptr && io->edge_index && io->cluster0, "structure_build: NULL input");
  DRGNN_REQUIRE(io->rowptr0 && io->col0 && io->eid0 && io->cscptr0 && io->cscrow0 && io->csceid0 && io->cl0 &&
                    io->cmptr0 && io->cmem0 && io->kptr0 && io->batch1 && io->rowptr1 && io->col1 && io->cscptr1 &&
                    io->cscrow1 && io->csceid1 && io->counts && io->status && io->gstat && io->scratch_n &&
                    io->scratch_e,
                "structure_build: NULL output / workspace");
  if (io->cluster1) {
    DRGNN_REQUIRE(io->c1_ptr && io->cl1 && io->cmptr1 && io->cmem1 && io->kptr1 && io->batch2,
                  "structure_build: cluster1 given but level-1 outputs are NULL");
  }
  if (io->ne > 0) DRGNN_REQUIRE(io->edge_attr != nullptr, "structure_build: ne > 0 but edge_attr is NULL");
  if (io->w0csr || io->w0csc || io->w1csc || io->edge_attr1) {
    DRGNN_REQUIRE(io->edge_attr != nullptr && io->ne >= 1, "structure_build: edge weights requested without edge_attr");
    DRGNN_REQUIRE(io->scratch_f != nullptr, "structure_build: scratch_f is NULL");
  }
  if (io->w1csc) DRGNN_REQUIRE(io->edge_attr1 != nullptr, "structure_build: w1csc needs edge_attr1");
  if (io->wblob && io->edge_attr)
    DRGNN_REQUIRE(io->blob != nullptr && io->edge_attr1 != nullptr && io->scratch_f != nullptr,
                  "structure_build: wblob needs blob, edge_attr1 and scratch_f");
  const int max_c1 = io->L1 > 0 ? io->max_n : 0;
  int64_t smem = drgnn_structure_smem_bytes(io->max_n, io->max_e, max_c1);
  if (smem < 0)
    return fail(DRGNN_ERR_UNSUPPORTED,
                "structure_build: a graph with %d nodes / %d edges does not fit one CTA's shared memory",
                io->max_n, io->max_e);
  static thread_local int64_t configured = -1;
  if (smem > configured) {
    DRGNN_CHECK_CUDA(cudaFuncSetAttribute(graph_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)device_info().smem_optin - kStaticSmemReserve));
    configured = device_info().smem_optin - kStaticSmemReserve;
  }
  cudaStream_t st = (cudaStream_t)stream;
  graph_local_kernel<<<io->B, kThreads, smem, st>>>(*io);
  DRGNN_CHECK_LAUNCH("graph_local_kernel");
  graph_finalize_kernel<<<io->B, 256, 0, st>>>(*io);
  DRGNN_CHECK_LAUNCH("graph_finalize_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_debug_structure_cycles(uint64_t* out32) {
  DRGNN_REQUIRE(out32 != nullptr, "debug_structure_cycles: NULL");
  DRGNN_CHECK_CUDA(cudaMemcpyFromSymbol(out32, g_sphase, sizeof(unsigned long long) * 32));
  return DRGNN_OK;
}

extern "C" int drgnn_cluster_offset(int64_t* cluster, const int32_t* seg_ptr, int32_t B, int64_t* work, void* stream) {
  DRGNN_REQUIRE(cluster && seg_ptr && work, "cluster_offset: NULL pointer");
  if (B <= 1) return DRGNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  seg_max_kernel<<<B, 256, 0, st>>>(cluster, seg_ptr, B, work);
  DRGNN_CHECK_LAUNCH("seg_max_kernel");
  seg_offset_add_kernel<<<B, 256, 0, st>>>(cluster, seg_ptr, B, work);
  DRGNN_CHECK_LAUNCH("seg_offset_add_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_ptr_from_sorted_ids(const int64_t* ids, int32_t n, int32_t B, int32_t* ptr, int32_t* status,
                                         void* stream) {
  DRGNN_REQUIRE(ptr && status && (ids || n == 0), "ptr_from_sorted_ids: NULL pointer");
  cudaStream_t st = (cudaStream_t)stream;
  ptr_from_sorted_ids_kernel<<<(n + 1 + 255) / 256, 256, 0, st>>>(ids, n, B, ptr, status);
  DRGNN_CHECK_LAUNCH("ptr_from_sorted_ids_kernel");
  return DRGNN_OK;
}
