// Per-graph fused GINet forward and backward: ONE CTA per protein-interface graph.
//
// The op-by-op path (aggregate / linear / maxpool / segment_mean launches, engine.py) spends a
// training step of the headline configuration (64 graphs of 200 nodes) in ~15 launches that are
// each a handful of dependent memory round trips: pure latency.  Graphs are independent and a
// whole graph (200 x 32 fp32 = 25.6 KB) fits shared memory, so these kernels keep
// every intermediate of a graph on the SM:
//
//   ginet_graph_fwd_kernel  (ginet.py:105-134, both branches at once)
//     x tile -> AX = A x -> Z1 = relu(AX W1cat^T) -> P1 = cluster max (+argmax)
//            -> AP = A1 P1 -> Z2 = relu(grouped AP W2^T) -> P2 = cluster max (+argmax) -> R = mean P2
//     what the backward needs goes to global memory once (AX, Z1, argmax0, AP, Z2, argmax1).
//   ginet_graph_bwd_kernel  (autograd of the above)
//     dR -> dZ2 (mean bwd, argmax routing, ReLU mask) -> per-graph dW2 partial, dAP = dZ2 W2
//        -> dP1 = A1^T dAP -> dZ1 (argmax routing, ReLU mask) -> per-graph dW1 partial
//   ginet_wgrad_reduce_kernel: sum of the per-graph partials in graph order (deterministic).
//
// Dense products use the CTA-level tile_gemm (8 x 4 register tiles over transposed shared-memory
// operands); gathers read neighbour rows from shared memory; indices come from the structure pass
// (L2 resident, just written).  GINet has no bias, no self term, unit edge weights (alpha == 1).
#include <cooperative_groups.h>
#include <float.h>
#include <string.h>

#include "common.cuh"

namespace drgnn {

static constexpr int FU_THREADS = 512;

// Diagnostic: SM clock at the phase boundaries of the CTA that runs graph 0 (drgnn_debug_phase_cycles).
__device__ unsigned long long g_phase[32];
#define DRGNN_PHASE(i)                                                         \
  do {                                                                         \
    if (blockIdx.x == 0 && threadIdx.x == 0) g_phase[i] = (unsigned long long)clock64(); \
  } while (0)

__host__ __device__ inline int up8(int x) { return (x + 7) & ~7; }

struct FwdSmem {
  int xs, axT, z1, w1t, w2t, p1, apT, z2, p2, total;
  int n_p, k_p, q_p;
};
__host__ __device__ inline FwdSmem fwd_plan(int F, int C1, int C2, int h1, int h2, int nb, int max_n, int max_k, int max_q) {
  FwdSmem p;
  p.n_p = up8(max_n) + 4;   // +4: transposing stores of consecutive rows spread over banks
  p.k_p = up8(max_k) + 4;
  p.q_p = up8(max_q);
  int o = 0;
  p.xs = o;  o += up8(max_n) * F;
  p.axT = o; o += F * p.n_p;
  p.z1 = o;  o += up8(max_n) * C1;
  p.w1t = o; o += F * C1;
  p.w2t = o; o += nb * h1 * h2;
  p.p1 = o;  o += up8(max_k) * C1;
  p.apT = o; o += C1 * p.k_p;
  p.z2 = o;  o += up8(max_k) * C2;
  p.p2 = o;  o += p.q_p * C2;
  p.total = o;
  return p;
}

struct BwdSmem {
  int dz2, dz2T, ap, dap, dp1, dz1, ax, w2, total;
  int k_p;
};
__host__ __device__ inline BwdSmem bwd_plan(int F, int C1, int C2, int h1, int h2, int nb, int max_n, int max_k) {
  BwdSmem p;
  p.k_p = up8(max_k) + 4;
  int o = 0;
  p.dz2 = o;  o += up8(max_k) * C2;
  p.dz2T = o; o += C2 * p.k_p;
  p.ap = o;   o += up8(max_k) * C1;
  p.dap = o;  o += up8(max_k) * C1;
  p.dp1 = o;  o += up8(max_k) * C1;
  p.dz1 = o;  o += up8(max_n) * C1;
  p.ax = o;   o += up8(max_n) * F;
  p.w2 = o;   o += nb * h2 * h1;
  p.total = o;
  return p;
}

// Forward of graph g out of the shared-memory workspace `fs`.  Returns false (after flagging) if the
// graph exceeds the host bounds.  `rrow` (shared, C2 floats) receives the graph's read-out row.
__device__ __forceinline__ bool graph_fwd_body(const drgnn_ginet_fused_args& a, float* fs, int g, float* rrow) {
  const int F = a.F, h1 = a.h1, h2 = a.h2, nb = a.nb;
  const int C1 = nb * h1, C2 = nb * h2;
  const FwdSmem P = fwd_plan(F, C1, C2, h1, h2, nb, a.max_n, a.max_k, a.max_q);
  float* xs = fs + P.xs;
  float* axT = fs + P.axT;
  float* z1 = fs + P.z1;
  float* w1t = fs + P.w1t;
  float* w2t = fs + P.w2t;
  float* p1 = fs + P.p1;
  float* apT = fs + P.apT;
  float* z2 = fs + P.z2;
  float* p2 = fs + P.p2;
  const int t = threadIdx.x, T = blockDim.x;
  const int n0 = a.node_ptr[g], n = a.node_ptr[g + 1] - n0;
  const int k0 = a.kptr0[g], K = a.kptr0[g + 1] - k0;
  const int q0 = a.kptr1[g], Q = a.kptr1[g + 1] - q0;
  if (n > a.max_n || K > a.max_k || Q > a.max_q) {  // host bounds violated: flag and leave (checked by validate())
    if (t == 0) atomicOr(a.status, 64);
    return false;
  }
  const int n8 = up8(n), K8 = up8(K);
  const int F4 = F >> 2, C14 = C1 >> 2, C24 = C2 >> 2;

  // ---- stage the graph's features and the (transposed) weights
  {
    const float4* src = reinterpret_cast<const float4*>(a.x + (int64_t)n0 * F);
    float4* dst = reinterpret_cast<float4*>(xs);
    for (int i = t; i < n * F4; i += T) dst[i] = src[i];
    for (int i = t; i < C1 * F; i += T) {      // W1 [C1][F] -> w1t [F][C1]
      const int c = i / F, f = i - c * F;
      w1t[f * C1 + c] = a.W1[i];
    }
    for (int i = t; i < nb * h2 * h1; i += T) {  // W2 [nb][h2][h1] -> w2t [nb][h1][h2]
      const int gg = i / (h2 * h1), r = i - gg * h2 * h1;
      const int o = r / h1, j = r - o * h1;
      w2t[(gg * h1 + j) * h2 + o] = a.W2[i];
    }
  }
  __syncthreads();
  DRGNN_PHASE(1);
  // ---- AX = A x   (thread per (row, 4 channels); row index fastest => conflict-free transposed stores)
  for (int item = t; item < n8 * F4; item += T) {
    const int q4 = item / n8, i = item - q4 * n8;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < n) {
      const int s = __ldg(a.rowptr0 + n0 + i), e = __ldg(a.rowptr0 + n0 + i + 1);
      for (int p = s; p < e; ++p) {
        const int c = __ldg(a.col0 + p) - n0;
        const float4 v = *reinterpret_cast<const float4*>(xs + c * F + q4 * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(a.Zin1 + (int64_t)(n0 + i) * F + q4 * 4) = acc;
    }
    axT[(q4 * 4 + 0) * P.n_p + i] = acc.x;
    axT[(q4 * 4 + 1) * P.n_p + i] = acc.y;
    axT[(q4 * 4 + 2) * P.n_p + i] = acc.z;
    axT[(q4 * 4 + 3) * P.n_p + i] = acc.w;
  }
  __syncthreads();
  DRGNN_PHASE(2);
  // ---- Z1 = relu(AX W1cat^T)
  tile_gemm(axT, P.n_p, w1t, C1, n8, C1, F, [&](int m, int c, float v) {
    v = v < 0.f ? 0.f : v;
    z1[m * C1 + c] = v;
    if (m < n) a.Z1[(int64_t)(n0 + m) * C1 + c] = v;
  });
  __syncthreads();
  DRGNN_PHASE(3);
  // ---- P1 = cluster max of Z1 (first member wins ties, a NaN never wins; community_pooling.py:201)
  for (int item = t; item < K * C14; item += T) {
    const int k = item / C14, q4 = item - k * C14;
    const int s = __ldg(a.cmptr0 + k0 + k), e = __ldg(a.cmptr0 + k0 + k + 1);
    float4 best = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    int4 arg = make_int4(-1, -1, -1, -1);
    for (int p = s; p < e; ++p) {
      const int i = __ldg(a.cmem0 + p);
      const float4 v = *reinterpret_cast<const float4*>(z1 + (i - n0) * C1 + q4 * 4);
      if (v.x > best.x) { best.x = v.x; arg.x = i; }
      if (v.y > best.y) { best.y = v.y; arg.y = i; }
      if (v.z > best.z) { best.z = v.z; arg.z = i; }
      if (v.w > best.w) { best.w = v.w; arg.w = i; }
    }
    if (arg.x < 0) best.x = 0.f;
    if (arg.y < 0) best.y = 0.f;
    if (arg.z < 0) best.z = 0.f;
    if (arg.w < 0) best.w = 0.f;
    *reinterpret_cast<float4*>(p1 + k * C1 + q4 * 4) = best;
    *reinterpret_cast<int4*>(a.arg0 + (int64_t)(k0 + k) * C1 + q4 * 4) = arg;
  }
  __syncthreads();
  DRGNN_PHASE(4);
  // ---- AP = A1 P1 on the coarsened graph
  for (int item = t; item < K8 * C14; item += T) {
    const int q4 = item / K8, k = item - q4 * K8;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (k < K) {
      const int s = __ldg(a.rowptr1 + k0 + k), e = __ldg(a.rowptr1 + k0 + k + 1);
      for (int p = s; p < e; ++p) {
        const int c = __ldg(a.col1 + p) - k0;
        const float4 v = *reinterpret_cast<const float4*>(p1 + c * C1 + q4 * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      *reinterpret_cast<float4*>(a.Zin2 + (int64_t)(k0 + k) * C1 + q4 * 4) = acc;
    }
    apT[(q4 * 4 + 0) * P.k_p + k] = acc.x;
    apT[(q4 * 4 + 1) * P.k_p + k] = acc.y;
    apT[(q4 * 4 + 2) * P.k_p + k] = acc.z;
    apT[(q4 * 4 + 3) * P.k_p + k] = acc.w;
  }
  __syncthreads();
  DRGNN_PHASE(5);
  // ---- Z2 = relu(AP_g W2_g^T) per branch g
  for (int gg = 0; gg < nb; ++gg) {
    tile_gemm(apT + gg * h1 * P.k_p, P.k_p, w2t + gg * h1 * h2, h2, K8, h2, h1, [&](int m, int o, float v) {
      v = v < 0.f ? 0.f : v;
      z2[m * C2 + gg * h2 + o] = v;
      if (m < K) a.Z2[(int64_t)(k0 + m) * C2 + gg * h2 + o] = v;
    });
  }
  __syncthreads();
  DRGNN_PHASE(6);
  // ---- P2 = level-1 cluster max (max_pool_x)
  for (int item = t; item < Q * C24; item += T) {
    const int q = item / C24, q4 = item - q * C24;
    const int s = __ldg(a.cmptr1 + q0 + q), e = __ldg(a.cmptr1 + q0 + q + 1);
    float4 best = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    int4 arg = make_int4(-1, -1, -1, -1);
    for (int p = s; p < e; ++p) {
      const int k = __ldg(a.cmem1 + p);
      const float4 v = *reinterpret_cast<const float4*>(z2 + (k - k0) * C2 + q4 * 4);
      if (v.x > best.x) { best.x = v.x; arg.x = k; }
      if (v.y > best.y) { best.y = v.y; arg.y = k; }
      if (v.z > best.z) { best.z = v.z; arg.z = k; }
      if (v.w > best.w) { best.w = v.w; arg.w = k; }
    }
    if (arg.x < 0) best.x = 0.f;
    if (arg.y < 0) best.y = 0.f;
    if (arg.z < 0) best.z = 0.f;
    if (arg.w < 0) best.w = 0.f;
    *reinterpret_cast<float4*>(p2 + q * C2 + q4 * 4) = best;
    *reinterpret_cast<int4*>(a.arg1 + (int64_t)(q0 + q) * C2 + q4 * 4) = arg;
  }
  __syncthreads();
  DRGNN_PHASE(7);
  // ---- R[g] = mean over the graph's level-1 clusters (scatter_mean by batch, ascending order)
  for (int c = t; c < C2; c += T) {
    float acc = 0.f;
    for (int q = 0; q < Q; ++q) acc += p2[q * C2 + c];
    acc *= 1.f / (float)max(Q, 1);
    a.R[(int64_t)g * C2 + c] = acc;
    if (rrow) rrow[c] = acc;
  }
  return true;
}

__global__ void __launch_bounds__(FU_THREADS, 1) ginet_graph_fwd_kernel(const drgnn_ginet_fused_args a) {
  extern __shared__ __align__(16) float fs[];
  graph_fwd_body(a, fs, blockIdx.x, nullptr);
}

// Backward of graph g: dR row (global or shared) -> per-graph partials of dW1 (`part1`, [C1][F]) and dW2
// (`part2`, [nb][h2][h1]).
__device__ __forceinline__ void graph_bwd_body(const drgnn_ginet_fused_args& a, float* bs, int g, const float* dRrow,
                                               float* part1, float* part2) {
  const int F = a.F, h1 = a.h1, h2 = a.h2, nb = a.nb;
  const int C1 = nb * h1, C2 = nb * h2;
  const BwdSmem P = bwd_plan(F, C1, C2, h1, h2, nb, a.max_n, a.max_k);
  float* dz2 = bs + P.dz2;
  float* dz2T = bs + P.dz2T;
  float* ap = bs + P.ap;
  float* dap = bs + P.dap;
  float* dp1 = bs + P.dp1;
  float* dz1 = bs + P.dz1;
  float* ax = bs + P.ax;
  float* w2 = bs + P.w2;
  const int t = threadIdx.x, T = blockDim.x;
  const int n0 = a.node_ptr[g], n = a.node_ptr[g + 1] - n0;
  const int k0 = a.kptr0[g], K = a.kptr0[g + 1] - k0;
  const int q0 = a.kptr1[g], Q = a.kptr1[g + 1] - q0;
  const int E1 = C1 * F, E2 = nb * h2 * h1;
  if (n > a.max_n || K > a.max_k) {
    for (int i = t; i < E1; i += T) part1[i] = 0.f;
    for (int i = t; i < E2; i += T) part2[i] = 0.f;
    return;
  }
  const int n8 = up8(n), K8 = up8(K);
  const float invQ = 1.f / (float)max(Q, 1);

  for (int i = t; i < E2; i += T) w2[i] = a.W2[i];
  // ---- dZ2: read-out mean backward, routed to the arg-max member, gated by ReLU; both layouts.
  // One item = 4 channels of one pooled node: the dependent chain cl1 -> arg1 is paid once per 16 bytes
  // and the unrolled loop keeps several items' loads in flight.
  {
    const int C24 = C2 >> 2;
#pragma unroll 2
    for (int item = t; item < K8 * C24; item += T) {
      const int k = item / C24, q4 = item - k * C24;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < K) {
        const int q = __ldg(a.cl1 + k0 + k);
        const int4 am = *reinterpret_cast<const int4*>(a.arg1 + (int64_t)q * C2 + q4 * 4);   // plain loads: written
        const float4 z = *reinterpret_cast<const float4*>(a.Z2 + (int64_t)(k0 + k) * C2 + q4 * 4);  // by this CTA
        const float4 d = *reinterpret_cast<const float4*>(dRrow + q4 * 4);
        const int me = k0 + k;
        v.x = (am.x == me && z.x > 0.f) ? d.x * invQ : 0.f;
        v.y = (am.y == me && z.y > 0.f) ? d.y * invQ : 0.f;
        v.z = (am.z == me && z.z > 0.f) ? d.z * invQ : 0.f;
        v.w = (am.w == me && z.w > 0.f) ? d.w * invQ : 0.f;
      }
      *reinterpret_cast<float4*>(dz2 + k * C2 + q4 * 4) = v;
      dz2T[(q4 * 4 + 0) * P.k_p + k] = v.x;
      dz2T[(q4 * 4 + 1) * P.k_p + k] = v.y;
      dz2T[(q4 * 4 + 2) * P.k_p + k] = v.z;
      dz2T[(q4 * 4 + 3) * P.k_p + k] = v.w;
    }
    const int C14v = C1 >> 2;
    const float4* src = reinterpret_cast<const float4*>(a.Zin2 + (int64_t)k0 * C1);
    float4* dst = reinterpret_cast<float4*>(ap);
    for (int item = t; item < K8 * C14v; item += T)
      dst[item] = (item / C14v) < K ? src[item] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  DRGNN_PHASE(12);
  // ---- per-graph dW2 partial [nb][h2][h1] = dZ2_g^T AP_g ; dAP = dZ2_g W2_g
  for (int gg = 0; gg < nb; ++gg) {
    tile_gemm(dz2 + gg * h2, C2, ap + gg * h1, C1, h2, h1, K8, [&](int o, int j, float v) {
      part2[(gg * h2 + o) * h1 + j] = v;
    });
    tile_gemm(dz2T + gg * h2 * P.k_p, P.k_p, w2 + gg * h2 * h1, h1, K8, h1, h2, [&](int m, int j, float v) {
      dap[m * C1 + gg * h1 + j] = v;
    });
  }
  __syncthreads();
  DRGNN_PHASE(13);
  // ---- dP1 = A1^T dAP  (CSC of the coarsened graph)
  const int C14 = C1 >> 2;
  for (int item = t; item < K * C14; item += T) {
    const int k = item / C14, q4 = item - k * C14;
    const int s = __ldg(a.cscptr1 + k0 + k), e = __ldg(a.cscptr1 + k0 + k + 1);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = s; p < e; ++p) {
      const int r = __ldg(a.cscrow1 + p) - k0;
      const float4 v = *reinterpret_cast<const float4*>(dap + r * C1 + q4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(dp1 + k * C1 + q4 * 4) = acc;
  }
  __syncthreads();
  DRGNN_PHASE(14);
  // ---- dZ1: routed to the arg-max node of its cluster, gated by ReLU (4 channels per item); stage AX
  {
#pragma unroll 4
    for (int item = t; item < n8 * C14; item += T) {
      const int i = item / C14, q4 = item - i * C14;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n) {
        const int k = __ldg(a.cl0 + n0 + i);
        const int4 am = *reinterpret_cast<const int4*>(a.arg0 + (int64_t)k * C1 + q4 * 4);
        const float4 z = *reinterpret_cast<const float4*>(a.Z1 + (int64_t)(n0 + i) * C1 + q4 * 4);
        const float4 d = *reinterpret_cast<const float4*>(dp1 + (k - k0) * C1 + q4 * 4);
        const int me = n0 + i;
        v.x = (am.x == me && z.x > 0.f) ? d.x : 0.f;
        v.y = (am.y == me && z.y > 0.f) ? d.y : 0.f;
        v.z = (am.z == me && z.z > 0.f) ? d.z : 0.f;
        v.w = (am.w == me && z.w > 0.f) ? d.w : 0.f;
      }
      *reinterpret_cast<float4*>(dz1 + i * C1 + q4 * 4) = v;
    }
    const int F4 = F >> 2;
    const float4* src = reinterpret_cast<const float4*>(a.Zin1 + (int64_t)n0 * F);
    float4* dst = reinterpret_cast<float4*>(ax);
#pragma unroll 4
    for (int item = t; item < n8 * F4; item += T) {
      float4 v = (item / F4) < n ? src[item] : make_float4(0.f, 0.f, 0.f, 0.f);
      v.x = (v.x == v.x) ? v.x : 0.f; v.y = (v.y == v.y) ? v.y : 0.f;
      v.z = (v.z == v.z) ? v.z : 0.f; v.w = (v.w == v.w) ? v.w : 0.f;
      dst[item] = v;
    }
  }
  __syncthreads();
  DRGNN_PHASE(15);
  // ---- per-graph dW1 partial [C1][F] = dZ1^T AX
  tile_gemm(dz1, C1, ax, F, C1, F, n8, [&](int c, int f, float v) { part1[c * F + f] = v; });
}

__global__ void __launch_bounds__(FU_THREADS, 1) ginet_graph_bwd_kernel(const drgnn_ginet_fused_args a) {
  extern __shared__ __align__(16) float bs[];
  const int g = blockIdx.x;
  const int E1 = a.nb * a.h1 * a.F, E2 = a.nb * a.h2 * a.h1;
  float* part = a.partial + (int64_t)g * (E1 + E2);
  graph_bwd_body(a, bs, g, a.dR + (int64_t)g * a.nb * a.h2, part, part + E1);
}

// ---------------------------------------------------------------------------------------
// Whole training step of one graph in ONE launch: forward, the network head on the graph's own
// read-out row (fc1 / ReLU / dropout / fc2 are row-wise, the loss is a sum of per-graph terms),
// and the backward down to per-graph gradient partials of EVERY parameter, laid out like the flat
// gradient buffer.  ginet_step_reduce_kernel then sums the partial rows in graph order.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float hash_uniform(uint32_t seed, uint32_t ctr, uint32_t idx) {
  uint32_t x = idx * 0x9E3779B1u ^ (ctr * 0x85EBCA77u) ^ (seed * 0xC2B2AE3Du);
  x ^= x >> 16; x *= 0x7FEB352Du;
  x ^= x >> 15; x *= 0x846CA68Bu;
  x ^= x >> 16;
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}

__global__ void __launch_bounds__(FU_THREADS, 1) ginet_graph_step_kernel(const drgnn_ginet_step_args s) {
  extern __shared__ __align__(16) float ws[];
  const drgnn_ginet_fused_args& a = s.g;
  const int g = blockIdx.x;
  const int t = threadIdx.x, T = blockDim.x, lane = t & 31, warp = t >> 5, nwarps = T >> 5;
  const int C2 = a.nb * a.h2, Hd = s.Hd, out = s.out;
  DRGNN_PHASE(0);
  // head scratch lives behind the graph workspace
  float* rrow = ws + s.head_off;         // [C2]  read-out row R[g]
  float* hrow = rrow + C2;               // [Hd]  hidden activation (after ReLU / dropout)
  float* dhrow = hrow + Hd;              // [Hd]
  float* drrow = dhrow + Hd;             // [C2]  dLoss / dR[g]
  float* prow = drrow + C2;              // [out] prediction, then dLoss / dpred
  float* part = s.partial + (int64_t)g * s.partial_ld;
  const bool ok = graph_fwd_body(a, ws, g, rrow);
  __syncthreads();
  DRGNN_PHASE(8);
  if (!ok) {
    if (!s.forward_only)
      for (int i = t; i < s.n_params + 1; i += T) part[i] = 0.f;
    return;
  }
  // ---- fc1: warp per hidden unit, lanes over the read-out channels
  for (int j = warp; j < Hd; j += nwarps) {
    float acc = 0.f;
    for (int c = lane; c < C2; c += 32) acc = fmaf(rrow[c], __ldg(s.fc1_w + (int64_t)j * C2 + c), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += s.fc1_b ? __ldg(s.fc1_b + j) : 0.f;
      acc = acc < 0.f ? 0.f : acc;
      if (s.keep) {
        acc = s.keep[(int64_t)g * Hd + j] > 0.f ? acc * s.keep_scale : 0.f;
      } else if (s.drop_p > 0.f) {
        // counter-based dropout mask: hash of (seed, optimiser step, graph, unit) -> uniform in [0,1)
        const uint32_t ctr = (uint32_t)s.step_dev[0];
        acc = hash_uniform(s.seed, ctr, (uint32_t)(g * Hd + j)) >= s.drop_p ? acc * s.keep_scale : 0.f;
      }
      hrow[j] = acc;
    }
  }
  __syncthreads();
  DRGNN_PHASE(9);
  // ---- fc2: warp per output
  for (int o = warp; o < out; o += nwarps) {
    float acc = 0.f;
    for (int j = lane; j < Hd; j += 32) acc = fmaf(hrow[j], __ldg(s.fc2_w + (int64_t)o * Hd + j), acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += s.fc2_b ? __ldg(s.fc2_b + o) : 0.f;
      prow[o] = acc;
      s.pred[(int64_t)g * out + o] = acc;
    }
  }
  __syncthreads();
  DRGNN_PHASE(10);
  if (s.forward_only || s.task == 0) return;
  // ---- loss term of this graph and dLoss/dpred (one thread)
  if (t == 0) {
    float lg = 0.f;
    if (s.task == 3) {
      float mx = prow[0];
      for (int c = 1; c < out; ++c) mx = fmaxf(mx, prow[c]);
      float se = 0.f;
      for (int c = 0; c < out; ++c) se += expf(prow[c] - mx);
      const float lse = mx + logf(se);
      const int tc = (int)s.y_class[g];
      const float w = s.class_w ? s.class_w[tc] : 1.f;
      lg = w * (lse - prow[tc]);
      for (int c = 0; c < out; ++c) prow[c] = w * (expf(prow[c] - lse) - (c == tc ? 1.f : 0.f)) * s.inv_norm;
    } else {
      for (int c = 0; c < out; ++c) {
        float p = prow[c], dp = 1.f;
        if (s.task == 2) {
          p = 1.f / (1.f + expf(-p));
          dp = p * (1.f - p);
        }
        const float d = p - s.y[(int64_t)g * out + c];
        lg += d * d;
        prow[c] = 2.f * d * s.inv_norm * dp;
      }
    }
    part[s.n_params] = lg * s.inv_norm;   // summed into the loss by the reduce kernel
  }
  __syncthreads();
  DRGNN_PHASE(11);
  // ---- head backward: fc2 partials, dh, fc1 partials, dR row
  for (int i = t; i < out * Hd; i += T) part[s.off_fc2w + i] = prow[i / Hd] * hrow[i % Hd];
  for (int o = t; o < out; o += T) part[s.off_fc2b + o] = prow[o];
  for (int j = t; j < Hd; j += T) {
    float acc = 0.f;
    for (int o = 0; o < out; ++o) acc = fmaf(prow[o], __ldg(s.fc2_w + (int64_t)o * Hd + j), acc);
    acc = hrow[j] > 0.f ? acc * s.keep_scale : 0.f;
    dhrow[j] = acc;
    part[s.off_fc1b + j] = acc;
  }
  __syncthreads();
  for (int i = t; i < Hd * C2; i += T) part[s.off_fc1w + i] = dhrow[i / C2] * rrow[i % C2];
  // dR[c] = sum_j dh[j] W1[j][c]: warps split the hidden units, lanes over channels, partial sums in shared memory
  {
    float* red = ws;  // the graph workspace is free again: [nwarps][C2]
    for (int c = lane; c < C2; c += 32) {
      float acc = 0.f;
      for (int j = warp; j < Hd; j += nwarps) acc = fmaf(dhrow[j], __ldg(s.fc1_w + (int64_t)j * C2 + c), acc);
      red[warp * C2 + c] = acc;
    }
    __syncthreads();
    for (int c = t; c < C2; c += T) {
      float acc = 0.f;
      for (int w = 0; w < nwarps; ++w) acc += red[w * C2 + c];
      drrow[c] = acc;
    }
  }
  __syncthreads();
  DRGNN_PHASE(11);
  graph_bwd_body(a, ws, g, drrow, part + s.off_w1, part + s.off_w2);
  __syncthreads();
  DRGNN_PHASE(16);
}

// grads[e] = sum over graphs of partial[g][e]; the slot behind the parameters is the loss.
// A block owns 32 consecutive elements; its 8 warps each sum one contiguous eighth of the graphs
// (ascending, all loads of a thread independent and in flight together), the eight partial sums are
// combined in ascending order: a fixed summation tree, deterministic for a given batch size.
// With fuse_adam the torch.optim.Adam update of element e follows in the same thread (single-GPU
// runs: no all-reduce sits between the two); the last block to finish bumps the step counter
// (ticket in step_dev[1]) so that no thread of this launch can observe the new value.
static constexpr int RED_SPLITS = 8;
__global__ void __launch_bounds__(32 * RED_SPLITS) ginet_step_reduce_kernel(const drgnn_ginet_step_args s) {
  __shared__ float sh[3];
  __shared__ float psum[RED_SPLITS][32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  const int B = s.g.B, n = s.n_params;
  if (s.fuse_adam && threadIdx.x == 0) {
    const float st = s.step_dev[0] + 1.f;
    sh[0] = st;
    sh[1] = adam_bias_correction(s.beta1, st);
    sh[2] = adam_bias_correction(s.beta2, st);
  }
  {
    const int gs = (B + RED_SPLITS - 1) / RED_SPLITS;
    const int g0 = q * gs, g1 = min(B, g0 + gs);
    float acc = 0.f;
    if (e <= n) {
      const float* src = s.partial + e;
      int g = g0;
      for (; g + 8 <= g1; g += 8) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = src[(int64_t)(g + u) * s.partial_ld];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u];
      }
      for (; g < g1; ++g) acc += src[(int64_t)g * s.partial_ld];
    }
    psum[q][lane] = acc;
  }
  __syncthreads();
  if (q == 0 && e <= n) {
    float acc = 0.f;
#pragma unroll
    for (int u = 0; u < RED_SPLITS; ++u) acc += psum[u][lane];
    if (e < n) {
      s.grads[e] = acc;
      if (s.fuse_adam) {
        float mi = s.adam_m[e], vi = s.adam_v[e];
        mi = mi + (acc - mi) * (1.f - s.beta1);
        vi = vi * s.beta2 + (1.f - s.beta2) * acc * acc;
        s.adam_m[e] = mi;
        s.adam_v[e] = vi;
        const float denom = sqrtf(vi) / sqrtf(sh[2]) + s.eps;
        s.adam_p[e] = s.adam_p[e] - (s.lr / sh[1]) * (mi / denom);
      }
    } else if (s.loss) {
      s.loss[0] = acc;
    }
  }
  if (!s.fuse_adam) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned* ticket = reinterpret_cast<unsigned*>(s.step_dev + 1);
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    if (is_last) {
      *ticket = 0u;
      s.step_dev[0] = sh[0];
    }
  }
}

// dW[e] = sum over graphs (ascending) of partial[g][e]; E = C1*F + nb*h2*h1 contiguous outputs
__global__ void __launch_bounds__(256) ginet_wgrad_reduce_kernel(const float* __restrict__ partial, int B, int E1, int E2,
                                                                 float* __restrict__ dW1, float* __restrict__ dW2) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  const int E = E1 + E2;
  if (e >= E) return;
  float s = 0.f;
#pragma unroll 8
  for (int g = 0; g < B; ++g) s += partial[(int64_t)g * E + e];
  if (e < E1) dW1[e] = s;
  else dW2[e - E1] = s;
}

#include "tc_tiles.cuh"
#include "fused_step2.cuh"

static inline bool fused_shapes_ok(const drgnn_ginet_fused_args& a) {
  const int C1 = a.nb * a.h1, C2 = a.nb * a.h2;
  return a.F % 4 == 0 && a.h1 % 4 == 0 && a.h2 % 8 == 0 && C1 % 8 == 0 && C2 % 4 == 0 && a.nb >= 1 && a.max_n > 0 &&
         a.max_k > 0 && a.max_q > 0;
}

}  // namespace drgnn

using namespace drgnn;

extern "C" int64_t drgnn_ginet_fused_smem_bytes(int32_t F, int32_t h1, int32_t h2, int32_t nb, int32_t max_n, int32_t max_k,
                                                int32_t max_q, int32_t backward) {
  if (F <= 0 || h1 <= 0 || h2 <= 0 || nb <= 0 || max_n <= 0 || max_k <= 0 || max_q <= 0) return DRGNN_ERR_INVALID;
  if (F % 4 || h1 % 4 || h2 % 8 || (nb * h1) % 8) return DRGNN_ERR_UNSUPPORTED;
  const int C1 = nb * h1, C2 = nb * h2;
  const int64_t bytes = 4 * (int64_t)(backward ? bwd_plan(F, C1, C2, h1, h2, nb, max_n, max_k).total
                                                : fwd_plan(F, C1, C2, h1, h2, nb, max_n, max_k, max_q).total);
  if (bytes > device_info().smem_optin - 2048) return DRGNN_ERR_UNSUPPORTED;
  return bytes;
}

static int check_fused(const drgnn_ginet_fused_args* a) {
  DRGNN_REQUIRE(a != nullptr, "ginet_fused: args is NULL");
  DRGNN_REQUIRE(a->B >= 0, "ginet_fused: negative batch");
  DRGNN_REQUIRE(fused_shapes_ok(*a), "ginet_fused: unsupported shape (F %% 4, h1 %% 4, h2 %% 8, nb*h1 %% 8)");
  DRGNN_REQUIRE(a->node_ptr && a->kptr0 && a->kptr1 && a->status, "ginet_fused: NULL structure pointer");
  return DRGNN_OK;
}

extern "C" int drgnn_ginet_fused_fwd(const drgnn_ginet_fused_args* a, void* stream) {
  int rc = check_fused(a);
  if (rc) return rc;
  DRGNN_REQUIRE(a->x && a->W1 && a->W2 && a->rowptr0 && a->col0 && a->rowptr1 && a->col1 && a->cmptr0 && a->cmem0 &&
                    a->cmptr1 && a->cmem1 && a->Zin1 && a->Z1 && a->arg0 && a->Zin2 && a->Z2 && a->arg1 && a->R,
                "ginet_fused_fwd: NULL pointer");
  if (a->B == 0) return DRGNN_OK;
  const int64_t smem = drgnn_ginet_fused_smem_bytes(a->F, a->h1, a->h2, a->nb, a->max_n, a->max_k, a->max_q, 0);
  if (smem < 0) return fail(DRGNN_ERR_UNSUPPORTED, "ginet_fused_fwd: a graph of %d nodes does not fit shared memory", a->max_n);
  static thread_local int64_t configured = -1;
  if (smem > configured) {
    DRGNN_CHECK_CUDA(cudaFuncSetAttribute(ginet_graph_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)device_info().smem_optin - 2048));
    configured = device_info().smem_optin - 2048;
  }
  ginet_graph_fwd_kernel<<<a->B, FU_THREADS, smem, (cudaStream_t)stream>>>(*a);
  DRGNN_CHECK_LAUNCH("ginet_graph_fwd_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_ginet_fused_bwd(const drgnn_ginet_fused_args* a, void* stream) {
  int rc = check_fused(a);
  if (rc) return rc;
  DRGNN_REQUIRE(a->W2 && a->cscptr1 && a->cscrow1 && a->cl0 && a->cl1 && a->Zin1 && a->Z1 && a->arg0 && a->Zin2 && a->Z2 &&
                    a->arg1 && a->dR && a->partial && a->dW1 && a->dW2,
                "ginet_fused_bwd: NULL pointer");
  if (a->B == 0) return DRGNN_OK;
  const int64_t smem = drgnn_ginet_fused_smem_bytes(a->F, a->h1, a->h2, a->nb, a->max_n, a->max_k, a->max_q, 1);
  if (smem < 0) return fail(DRGNN_ERR_UNSUPPORTED, "ginet_fused_bwd: a graph of %d nodes does not fit shared memory", a->max_n);
  static thread_local int64_t configured = -1;
  if (smem > configured) {
    DRGNN_CHECK_CUDA(cudaFuncSetAttribute(ginet_graph_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)device_info().smem_optin - 2048));
    configured = device_info().smem_optin - 2048;
  }
  cudaStream_t st = (cudaStream_t)stream;
  ginet_graph_bwd_kernel<<<a->B, FU_THREADS, smem, st>>>(*a);
  DRGNN_CHECK_LAUNCH("ginet_graph_bwd_kernel");
  const int E1 = a->nb * a->h1 * a->F, E2 = a->nb * a->h2 * a->h1;
  ginet_wgrad_reduce_kernel<<<(E1 + E2 + 255) / 256, 256, 0, st>>>(a->partial, a->B, E1, E2, a->dW1, a->dW2);
  DRGNN_CHECK_LAUNCH("ginet_wgrad_reduce_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_debug_phase_cycles(uint64_t* out32) {
  DRGNN_REQUIRE(out32 != nullptr, "debug_phase_cycles: NULL");
  DRGNN_CHECK_CUDA(cudaMemcpyFromSymbol(out32, g_phase, sizeof(unsigned long long) * 32));
  return DRGNN_OK;
}

extern "C" int64_t drgnn_ginet_step_smem_bytes(int32_t F, int32_t h1, int32_t h2, int32_t nb, int32_t max_n, int32_t max_k,
                                               int32_t max_q, int32_t Hd, int32_t out) {
  const int64_t f = drgnn_ginet_fused_smem_bytes(F, h1, h2, nb, max_n, max_k, max_q, 0);
  const int64_t b = drgnn_ginet_fused_smem_bytes(F, h1, h2, nb, max_n, max_k, max_q, 1);
  if (f < 0) return f;
  if (b < 0) return b;
  if (Hd <= 0 || out <= 0) return DRGNN_ERR_INVALID;
  const int64_t head = 4 * (int64_t)(2 * nb * h2 + 2 * Hd + out + 8);
  const int64_t red = 4 * (int64_t)(FU_THREADS / 32) * nb * h2;
  int64_t bytes = (f > b ? f : b);
  if (bytes < red) bytes = red;
  bytes = ((bytes + 15) & ~(int64_t)15) + head;
  if (bytes > device_info().smem_optin - 2048) return DRGNN_ERR_UNSUPPORTED;
  return bytes;
}

extern "C" int64_t drgnn_ginet_step2_smem_bytes(int32_t F, int32_t h1, int32_t h2, int32_t max_n, int32_t max_k, int32_t max_q,
                                                int32_t max_e, int32_t Hd, int32_t out) {
  if (F <= 0 || h1 <= 0 || h2 <= 0 || max_n <= 0 || max_k <= 0 || max_q <= 0 || max_e <= 0 || Hd <= 0 || out <= 0)
    return DRGNN_ERR_INVALID;
  if (F % 4 || h1 % 4 || h2 % 4 || (2 * h2 * Hd) % 4) return DRGNN_ERR_UNSUPPORTED;
  // keep the word count far from int overflow before planning
  if ((int64_t)max_n * (F + h1) > (1 << 24) || max_e > (1 << 24) || (int64_t)Hd * h2 > (1 << 22)) return DRGNN_ERR_UNSUPPORTED;
  const int64_t bytes = 4 * (int64_t)step2_plan(F, h1, h2, max_n, max_k, max_q, max_e, Hd, out).total;
  if (bytes > device_info().smem_optin - 1024) return DRGNN_ERR_UNSUPPORTED;
  return bytes;
}

// How many 2-CTA clusters of ginet_graph_step2_kernel the device holds at once (cached per size).
static int step2_max_clusters(int64_t smem) {
  static thread_local int64_t cached_smem = -1;
  static thread_local int cached = 0;
  if (cached_smem == smem) return cached;
  int best = 0;
  for (int with_attr = 0; with_attr < 2 && best <= 0; ++with_attr) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * (unsigned)device_info().sms);
    cfg.blockDim = dim3(S2_THREADS);
    cfg.dynamicSmemBytes = (size_t)smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = with_attr ? attr : nullptr;
    cfg.numAttrs = with_attr ? 1 : 0;
    int nc = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, ginet_graph_step2_kernel, &cfg);
    if (e != cudaSuccess) {
      fail(DRGNN_ERR_CUDA, "cudaOccupancyMaxActiveClusters(%s): %s", with_attr ? "attr" : "compile-time dims", cudaGetErrorString(e));
      (void)cudaGetLastError();
      nc = 0;
    }
    if (nc > best) best = nc;
  }
  cached = best;
  cached_smem = smem;
  return best;
}
extern "C" int drgnn_ginet_step2_max_clusters(int64_t smem_bytes) {
  if (smem_bytes < 0 || smem_bytes > device_info().smem_optin - 1024) return DRGNN_ERR_INVALID;
  if (cudaFuncSetAttribute(ginet_graph_step2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)device_info().smem_optin - 1024) != cudaSuccess) {
    (void)cudaGetLastError();
    return DRGNN_ERR_CUDA;
  }
  return step2_max_clusters(smem_bytes);
}

static thread_local int g_step_variant = 0;
static thread_local int g_step_launches = 0;
extern "C" int drgnn_ginet_step_last_variant(void) { return g_step_variant; }
extern "C" int drgnn_ginet_step_last_launches(void) { return g_step_launches; }

extern "C" int drgnn_ginet_step(const drgnn_ginet_step_args* s, void* stream) {
  DRGNN_REQUIRE(s != nullptr, "ginet_step: args is NULL");
  const drgnn_ginet_fused_args* a = &s->g;
  int rc = check_fused(a);
  if (rc) return rc;
  DRGNN_REQUIRE(a->x && a->W1 && a->W2 && a->rowptr0 && a->col0 && a->rowptr1 && a->col1 && a->cmptr0 && a->cmem0 &&
                    a->cmptr1 && a->cmem1 && a->cscptr1 && a->cscrow1 && a->cl0 && a->cl1 && a->Zin1 && a->Z1 && a->arg0 &&
                    a->Zin2 && a->Z2 && a->arg1 && a->R,
                "ginet_step: NULL graph pointer");
  DRGNN_REQUIRE(s->fc1_w && s->fc2_w && s->pred && s->Hd > 0 && s->out > 0, "ginet_step: NULL / bad head arguments");
  DRGNN_REQUIRE(s->task >= 0 && s->task <= 3, "ginet_step: bad task %d", s->task);
  if (!s->forward_only && s->task != 0) {
    DRGNN_REQUIRE(s->partial && s->grads && s->n_params > 0 && s->partial_ld > s->n_params, "ginet_step: bad gradient buffers");
    DRGNN_REQUIRE(s->task == 3 ? (s->y_class != nullptr) : (s->y != nullptr), "ginet_step: missing targets");
    DRGNN_REQUIRE(!s->fuse_adam || (s->adam_p && s->adam_m && s->adam_v && s->step_dev), "ginet_step: fuse_adam needs the Adam buffers");
  }
  DRGNN_REQUIRE(s->keep || s->drop_p <= 0.f || s->step_dev, "ginet_step: hashed dropout needs step_dev");
  DRGNN_REQUIRE(s->variant >= 0 && s->variant <= 2, "ginet_step: bad variant %d", s->variant);
  if (a->B == 0) return DRGNN_OK;
  cudaStream_t st = (cudaStream_t)stream;
  // ---- cluster kernel: a pair of CTAs per graph, one branch each, everything in shared memory
  int64_t smem2 = -1;
  if (s->variant != 1 && step2_shapes_ok(*s) && ((uintptr_t)a->x % 16) == 0 && ((uintptr_t)s->fc1_w % 16) == 0)
    smem2 = drgnn_ginet_step2_smem_bytes(a->F, a->h1, a->h2, a->max_n, a->max_k, a->max_q, s->max_e, s->Hd, s->out);
  if (s->variant == 2 && smem2 < 0)
    return fail(DRGNN_ERR_UNSUPPORTED, "ginet_step: the cluster kernel does not support this shape (nb %d, max_n %d, max_e %d)",
                a->nb, a->max_n, s->max_e);
  if (smem2 >= 0) {
    static thread_local int64_t configured2 = -1;
    if (smem2 > configured2) {
      DRGNN_CHECK_CUDA(cudaFuncSetAttribute(ginet_graph_step2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)device_info().smem_optin - 1024));
      configured2 = device_info().smem_optin - 1024;
    }
    Step2Plan plan = step2_plan(a->F, a->h1, a->h2, a->max_n, a->max_k, a->max_q, s->max_e, s->Hd, s->out);
    const bool train = !s->forward_only && s->task != 0;
    // One CTA per SM (shared memory): with 2B <= SMs the whole grid is co-resident and the gradient
    // reduction (+ Adam) can follow a grid barrier inside the same launch.  flags bit 1 disables it.
    const int G = 2 * a->B;
    const int occ_clusters = step2_max_clusters(smem2);   // co-resident 2-CTA clusters at this shared-memory size
    plan.fused_reduce = (train && !s->skip_reduce && s->step_dev != nullptr && !(s->flags & 2) && a->B <= occ_clusters) ? 1 : 0;
    drgnn_peer_comm comm;
    memset(&comm, 0, sizeof(comm));
    if (s->comm != nullptr && train && !s->skip_reduce) {
      // the exchange over peer memory runs inside the launch: every rank must take this very path
      const drgnn_peer_comm* c = s->comm;
      DRGNN_REQUIRE(c->world >= 1 && c->world <= DRGNN_MAX_PEERS && c->rank >= 0 && c->rank < c->world,
                    "ginet_step: bad world / rank %d / %d", c->world, c->rank);
      if (c->world > 1) {
        if (!plan.fused_reduce)
          return fail(DRGNN_ERR_UNSUPPORTED, "ginet_step: the in-kernel exchange needs the co-resident grid (B %d, clusters %d)",
                      a->B, occ_clusters);
        DRGNN_REQUIRE(s->fuse_adam, "ginet_step: the in-kernel exchange applies Adam (fuse_adam)");
        DRGNN_REQUIRE(c->ctr && c->max_blocks >= G && c->stride >= s->n_params + 1, "ginet_step: exchange layout too small");
        for (int r = 0; r < c->world; ++r)
          DRGNN_REQUIRE(c->xll[r], "ginet_step: rank %d has no low-latency exchange buffer", r);
        comm = *c;
      }
    }
    DRGNN_REQUIRE(((uintptr_t)s->zin1 % 16) == 0, "ginet_step: zin1 must be 16-byte aligned");
    drgnn_ginet_step_args k2 = *s;
    if (a->F % 8 || a->h1 % 8 || a->h2 % 8) k2.flags &= ~4;   // the tensor-core tiles need widths that are multiples of 8
    ginet_graph_step2_kernel<<<G, S2_THREADS, smem2, st>>>(k2, plan, comm);
    DRGNN_CHECK_LAUNCH("ginet_graph_step2_kernel");
    g_step_variant = 2;
    g_step_launches = plan.fused_reduce ? 1 : ((train && !s->skip_reduce) ? 2 : 1);
    if (train && !s->skip_reduce && !plan.fused_reduce) {
      ginet_step_reduce_kernel<<<(s->n_params + 1 + 31) / 32, 32 * RED_SPLITS, 0, st>>>(*s);
      DRGNN_CHECK_LAUNCH("ginet_step_reduce_kernel");
    }
    return DRGNN_OK;
  }
  g_step_variant = 1;
  g_step_launches = (!s->forward_only && s->task != 0 && !s->skip_reduce) ? 2 : 1;
  const int64_t smem = drgnn_ginet_step_smem_bytes(a->F, a->h1, a->h2, a->nb, a->max_n, a->max_k, a->max_q, s->Hd, s->out);
  if (smem < 0) return fail(DRGNN_ERR_UNSUPPORTED, "ginet_step: a graph of %d nodes does not fit shared memory", a->max_n);
  static thread_local int64_t configured = -1;
  if (smem > configured) {
    DRGNN_CHECK_CUDA(cudaFuncSetAttribute(ginet_graph_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)device_info().smem_optin - 2048));
    configured = device_info().smem_optin - 2048;
  }
  drgnn_ginet_step_args k = *s;
  const int64_t head = 4 * (int64_t)(2 * a->nb * a->h2 + 2 * s->Hd + s->out + 8);
  k.head_off = (int32_t)((smem - head) / 4);
  ginet_graph_step_kernel<<<a->B, FU_THREADS, smem, st>>>(k);
  DRGNN_CHECK_LAUNCH("ginet_graph_step_kernel");
  if (!s->forward_only && s->task != 0 && !s->skip_reduce) {
    ginet_step_reduce_kernel<<<(s->n_params + 1 + 31) / 32, 32 * RED_SPLITS, 0, st>>>(k);
    DRGNN_CHECK_LAUNCH("ginet_step_reduce_kernel");
  }
  return DRGNN_OK;
}
