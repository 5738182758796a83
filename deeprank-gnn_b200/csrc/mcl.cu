// GPU pre-clustering: Markov clustering of every graph of a batch, one CTA per graph.
//
// Replaces community_detection(edge_index, num_nodes, method='mcl') (deeprank_gnn/community_pooling.py:95-158),
// which PreCluster runs twice per graph on the CPU (DataSet.py:45-88) through networkx -> scipy -> the third-party
// package markov_clustering (run_mcl with its default parameters + get_clusters).  Same algorithm, same float64
// arithmetic, same labelling (clusters sorted as tuples, a node in several clusters keeps the last):
//
//   M <- adjacency (unit weights, undirected) with M[i,i] = 1, columns normalised to sum 1
//   repeat <= 100 times:  E = M M ; E = E .* E ; normalise columns ; drop entries < 0.001 but keep each column's
//                         maximum ; stop when |M_new - M| <= 1e-8 + 1e-5 |M| everywhere
//   attractors = rows with a non-zero diagonal ; cluster(a) = columns with a non-zero entry in row a
//   clusters sorted lexicographically as sorted tuples, duplicates merged ; label[v] = index of the LAST cluster
//   that contains v (0 if none)
//
// Every step of an iteration is column-local, so a WARP owns a column: it accumulates column j of M M in a
// shared-memory vector (lanes over the rows, M column-major in an L2-resident workspace: coalesced reads of
// M[:,k], broadcast of M[k,j], zero entries of column j skipped - the sums run over the non-zero k in ascending
// order like the sparse product of the reference), squares, normalises, prunes and compares it in place.
// The dense matrices make this an offline-quality kernel (n^2 doubles per graph), which is what pre-clustering
// is: it runs once per data set, not per step.
#include <float.h>

#include "common.cuh"

namespace drgnn {

static constexpr int MCL_THREADS = 512;
static constexpr int MCL_WARPS = MCL_THREADS / 32;

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// set A < set B in the order of their sorted member tuples?  (bitmasks over W words)
__device__ __forceinline__ int mcl_tuple_less(const uint32_t* a, const uint32_t* b, int W) {
  for (int w = 0; w < W; ++w) {
    const uint32_t x = a[w] ^ b[w];
    if (x) {
      const int d = __ffs(x) - 1;                 // lowest member the sets disagree on
      const bool in_a = (a[w] >> d) & 1u;
      const uint32_t* other = in_a ? b : a;       // the set WITHOUT d: smaller iff it has no member above d (prefix)
      bool above = (other[w] >> d) >> 1 != 0u;
      for (int v = w + 1; v < W && !above; ++v) above = other[v] != 0u;
      // the set with d is smaller iff the other set continues with a larger member
      return in_a ? (above ? 1 : 0) : (above ? 0 : 1);
    }
  }
  return 0;   // equal
}

__global__ void __launch_bounds__(MCL_THREADS, 1)
    mcl_graph_kernel(const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ edge_ptr, const void* __restrict__ edge_index,
                     int64_t E, int idx32, int max_n, double* __restrict__ work, int64_t* __restrict__ cluster, int32_t* __restrict__ iters,
                     int32_t* __restrict__ status, int max_iter, double threshold) {
  extern __shared__ __align__(16) double colacc[];          // [MCL_WARPS][max_n] doubles; reused as bitmasks at the end
  __shared__ int s_flag;
  const int g = blockIdx.x, t = threadIdx.x, lane = t & 31, w = t >> 5;
  const int n0 = node_ptr[g], n = node_ptr[g + 1] - n0;
  const int64_t e0 = edge_ptr[g], m = edge_ptr[g + 1] - e0;
  if (n <= 0) return;
  if (n > max_n) {
    if (t == 0) atomicOr(status, DRGNN_ST_FUSED_BOUNDS);
    return;
  }
  double* A = work + (int64_t)2 * g * max_n * max_n;         // column-major n x n
  double* Bm = A + (int64_t)max_n * max_n;
  const int64_t nn = (int64_t)n * n;
  for (int64_t i = t; i < nn; i += MCL_THREADS) A[i] = 0.0;
  __syncthreads();
  for (int64_t e = t; e < m; e += MCL_THREADS) {
    long long r, c;
    if (idx32) {
      r = reinterpret_cast<const int32_t*>(edge_index)[e0 + e] - n0;
      c = reinterpret_cast<const int32_t*>(edge_index)[E + e0 + e] - n0;
    } else {
      r = reinterpret_cast<const int64_t*>(edge_index)[e0 + e] - n0;
      c = reinterpret_cast<const int64_t*>(edge_index)[E + e0 + e] - n0;
    }
    if (r < 0 || r >= n || c < 0 || c >= n) {
      atomicOr(status, DRGNN_ST_EDGE_OUTSIDE_GRAPH);
      continue;
    }
    A[r + (int64_t)n * c] = 1.0;                              // nx.Graph: undirected, duplicates collapse
    A[c + (int64_t)n * r] = 1.0;
  }
  __syncthreads();
  for (int i = t; i < n; i += MCL_THREADS) A[i + (int64_t)n * i] = 1.0;   // add_self_loops(loop_value = 1)
  __syncthreads();
  for (int j = w; j < n; j += MCL_WARPS) {                    // normalize(norm='l1', axis=0)
    double* col = A + (int64_t)n * j;
    double s = 0.0;
    for (int i = lane; i < n; i += 32) s += fabs(col[i]);
    s = warp_sum_d(s);
    if (s == 0.0) s = 1.0;
    for (int i = lane; i < n; i += 32) col[i] = col[i] / s;
  }
  __syncthreads();
  double* acc = colacc + (size_t)w * max_n;
  int it = 0;
  for (it = 1; it <= max_iter; ++it) {
    if (t == 0) s_flag = 1;
    __syncthreads();
    int conv = 1;
    for (int j = w; j < n; j += MCL_WARPS) {
      const double* mj = A + (int64_t)n * j;
      for (int i = lane; i < n; i += 32) acc[i] = 0.0;
      // expansion: column j of M M, over the non-zero entries of column j in ascending k
      for (int k0 = 0; k0 < n; k0 += 32) {
        const double mine = (k0 + lane < n) ? mj[k0 + lane] : 0.0;
        unsigned nz = __ballot_sync(0xffffffffu, mine != 0.0);
        while (nz) {
          const int kl = __ffs(nz) - 1;
          nz &= nz - 1;
          const double v = __shfl_sync(0xffffffffu, mine, kl);
          const double* mk = A + (int64_t)n * (k0 + kl);
          for (int i = lane; i < n; i += 32) acc[i] = __dadd_rn(acc[i], __dmul_rn(mk[i], v));
        }
      }
      // inflation (power 2) + column normalisation
      double s = 0.0;
      for (int i = lane; i < n; i += 32) {
        const double v = acc[i] * acc[i];
        acc[i] = v;
        s += v;
      }
      s = warp_sum_d(s);
      if (s == 0.0) s = 1.0;
      // the column's maximum (first occurrence) survives the pruning whatever its value
      double best = -1.0;
      int bi = n;
      for (int i = lane; i < n; i += 32) {
        const double v = acc[i] / s;
        acc[i] = v;
        if (v > best) { best = v; bi = i; }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, best, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
      }
      double* out = Bm + (int64_t)n * j;
      for (int i = lane; i < n; i += 32) {
        double v = acc[i];
        if (v < threshold && i != bi) v = 0.0;
        const double last = mj[i];
        if (fabs(v - last) - 1e-5 * fabs(last) > 1e-8) conv = 0;
        out[i] = v;
      }
    }
    if (!conv) s_flag = 0;
    __syncthreads();
    double* tmp = A; A = Bm; Bm = tmp;
    const int done = s_flag;
    __syncthreads();
    if (done) break;
  }
  if (t == 0 && iters) iters[g] = it > max_iter ? max_iter : it;
  // ---- get_clusters + the labelling of community_detection
  const int W = (n + 31) >> 5;
  uint32_t* mask = reinterpret_cast<uint32_t*>(colacc);        // [n][W] bitmask of row a's non-zero columns
  int* rank = reinterpret_cast<int*>(mask + (size_t)n * W);    // [n]: -1 = not an attractor
  int* rep = rank + n;                                         // [n]
  for (int i = t; i < n * W; i += MCL_THREADS) mask[i] = 0u;
  __syncthreads();
  for (int64_t i = t; i < nn; i += MCL_THREADS) {
    const int a = (int)(i % n), c = (int)(i / n);
    if (A[i] != 0.0 && A[a + (int64_t)n * a] != 0.0) atomicOr(&mask[(size_t)a * W + (c >> 5)], 1u << (c & 31));
  }
  __syncthreads();
  // rep[b] = 1: b is the FIRST attractor that carries its cluster (duplicates are merged: set() in get_clusters)
  for (int b = t; b < n; b += MCL_THREADS) {
    int r = 0;
    if (A[b + (int64_t)n * b] != 0.0) {
      r = 1;
      const uint32_t* mb = mask + (size_t)b * W;
      for (int c2 = 0; c2 < b && r; ++c2) {
        if (A[c2 + (int64_t)n * c2] == 0.0) continue;
        const uint32_t* mc = mask + (size_t)c2 * W;
        bool eq = true;
        for (int ww = 0; ww < W && eq; ++ww) eq = mc[ww] == mb[ww];
        if (eq) r = 0;
      }
    }
    rep[b] = r;
  }
  __syncthreads();
  // rank[a] = number of distinct clusters that sort before cluster(a) = its index in sorted(set(clusters))
  for (int a = t; a < n; a += MCL_THREADS) {
    int rk = -1;
    if (A[a + (int64_t)n * a] != 0.0) {
      const uint32_t* ma = mask + (size_t)a * W;
      rk = 0;
      for (int b = 0; b < n; ++b)
        if (rep[b] && b != a && mcl_tuple_less(mask + (size_t)b * W, ma, W)) ++rk;
    }
    rank[a] = rk;
  }
  __syncthreads();
  for (int v = t; v < n; v += MCL_THREADS) {
    int lab = 0;
    for (int a = 0; a < n; ++a)
      if (rank[a] >= 0 && ((mask[(size_t)a * W + (v >> 5)] >> (v & 31)) & 1u)) lab = max(lab, rank[a]);
    cluster[n0 + v] = lab;
  }
}

}  // namespace drgnn

using namespace drgnn;

extern "C" int64_t drgnn_mcl_work_doubles(int32_t B, int32_t max_n) {
  if (B < 0 || max_n < 0) return DRGNN_ERR_INVALID;
  return (int64_t)2 * B * max_n * max_n;
}

extern "C" int drgnn_mcl_cluster(const int32_t* node_ptr, const int32_t* edge_ptr, const void* edge_index, int32_t B, int64_t E,
                                 int32_t idx32, int32_t max_n, double* work, int64_t* cluster, int32_t* iters, int32_t* status,
                                 void* stream) {
  DRGNN_REQUIRE(B >= 0 && E >= 0 && max_n >= 0, "mcl_cluster: negative size");
  if (B == 0 || max_n == 0) return DRGNN_OK;
  DRGNN_REQUIRE(node_ptr && edge_ptr && work && cluster && status, "mcl_cluster: NULL pointer");
  DRGNN_REQUIRE(E == 0 || edge_index, "mcl_cluster: NULL edge_index");
  // shared memory: the per-warp column accumulators, later the attractor bitmasks + ranks
  const int64_t W = (max_n + 31) / 32;
  int64_t smem = (int64_t)MCL_WARPS * max_n * 8;
  const int64_t need2 = (int64_t)max_n * W * 4 + (int64_t)max_n * 8 + 16;
  if (need2 > smem) smem = need2;
  if (smem > device_info().smem_optin - 1024)
    return fail(DRGNN_ERR_UNSUPPORTED, "mcl_cluster: graphs of %d nodes exceed the shared memory of one CTA", max_n);
  static thread_local int64_t configured = -1;
  if (smem > configured) {
    DRGNN_CHECK_CUDA(cudaFuncSetAttribute(mcl_graph_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)device_info().smem_optin - 1024));
    configured = device_info().smem_optin - 1024;
  }
  mcl_graph_kernel<<<B, MCL_THREADS, smem, (cudaStream_t)stream>>>(node_ptr, edge_ptr, edge_index, E, idx32, max_n, work, cluster,
                                                                   iters, status, 100, 0.001);
  DRGNN_CHECK_LAUNCH("mcl_graph_kernel");
  return DRGNN_OK;
}
