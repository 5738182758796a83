// Whole GINet training step of one graph on a PAIR of CTAs (thread-block cluster of 2), everything
// of the graph resident in shared memory.  Included by fused.cu (inside namespace drgnn).
//
// GINet (ginet.py:99-141) runs two structurally identical branches that only meet at the
// concatenated read-out row.  ginet_graph_step_kernel (v1) computes both branches in one CTA and
// re-reads indices / intermediates from global memory between its phases: every phase is a chain
// of dependent L2 round trips and only 64 of the 148 SMs work at batch 64.  Here
//
//   * the cluster's CTA r owns branch r (conv1_r -> pool -> conv2_r -> pool -> mean), so a batch of
//     64 graphs occupies 128 SMs and every dense product is half as wide;
//   * ONE asynchronous staging pass (cp.async, two commit groups) brings the graph's feature tile,
//     EVERY index slice of both graph levels (CSR, CSC, cluster members, cluster ids) and the head
//     weights into shared memory; after it no phase touches global memory except to store results;
//   * every intermediate the backward needs (AX, Z1, argmax0, AP, Z2, argmax1) stays in shared
//     memory (optionally mirrored to global memory for the parity tests, flag bit 0);
//   * the two halves of the read-out row are exchanged through distributed shared memory
//     (one cluster barrier); the tiny head (fc1 / ReLU / dropout / fc2 / loss) is evaluated by both
//     CTAs, its gradient rows are written half by each;
//   * the weight-gradient products (dW1 = dZ1^T AX, dW2 = dZ2^T AP: M, N tiny, K = nodes) are
//     split over K across the whole CTA and reduced in a fixed order (deterministic).
//
// Arithmetic per output element is the same fmaf chain, in the same order, as v1 / the op-level
// kernels for everything but the split-K weight gradients.
#pragma once
// (fused.cu includes <cooperative_groups.h> at global scope before entering the namespace)

namespace cgx = cooperative_groups;

static constexpr int S2_THREADS = 512;
static constexpr int S2_MAX_KS = 16;   // split-K factor bound of the weight-gradient products

__device__ __forceinline__ uint32_t s2_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s2_cp4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s2_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void s2_cp16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s2_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void s2_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void s2_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__host__ __device__ inline int s2_up4(int x) { return (x + 3) & ~3; }

// Shared-memory plan (offsets in 4-byte words, every offset a multiple of 4 = 16 bytes).
struct Step2Plan {
  // float regions
  int xs, ax, z1, dz1, w1t, w2t, w2, p1, ap, z2, p2, dz2, dap, dp1;
  int fc1w, fc2w, fc1b, fc2b, rrow, hrow, dhrow, drrow, prow, red;
  // int regions
  int arg0, arg1, rp0, col0, rp1, col1, cmp0, cmem0, cl0, cmp1, cmem1, cl1, cscp1, cscr1;
  int ldx, ldz1, ldp, ldz2;   // row strides (words): F+4, h1+4, h1+4, h2+4
  int xs_words;               // capacity of the xs region (split-K scratch in the backward)
  int max_e1;
  int total;
};
__host__ __device__ inline Step2Plan step2_plan(int F, int h1, int h2, int max_n, int max_k, int max_q, int max_e, int Hd,
                                                int out) {
  Step2Plan p;
  const int n8 = up8(max_n), k8 = up8(max_k), q8 = up8(max_q);
  const int C2 = 2 * h2;
  const long long e1 = (long long)max_k * (max_k - 1);
  p.max_e1 = (int)(e1 < (long long)max_e ? e1 : (long long)max_e);
  if (p.max_e1 < 1) p.max_e1 = 1;
  p.ldx = F + 4; p.ldz1 = h1 + 4; p.ldp = h1 + 4; p.ldz2 = h2 + 4;
  int o = 0;
  auto take = [&](int words) { const int at = o; o += s2_up4(words); return at; };
  int xsw = n8 * F;
  if (xsw < h1 * F) xsw = h1 * F;          // room for at least one split of each weight-gradient product
  if (xsw < h2 * h1) xsw = h2 * h1;
  p.xs_words = s2_up4(xsw);
  p.xs = take(xsw);
  p.ax = take(n8 * p.ldx);
  p.z1 = take(n8 * p.ldz1);
  p.dz1 = take(n8 * p.ldz1);
  p.w1t = take(F * h1);
  p.w2t = take(h1 * h2);
  p.w2 = take(h2 * h1);
  p.p1 = take(k8 * p.ldp);
  p.ap = take(k8 * p.ldp);
  p.z2 = take(k8 * p.ldz2);
  p.p2 = take(q8 * h2);
  p.dz2 = take(k8 * p.ldz2);
  p.dap = take(k8 * p.ldp);
  p.dp1 = take(k8 * p.ldp);
  p.fc1w = take(Hd * C2);
  p.fc2w = take(out * Hd);
  p.fc1b = take(Hd);
  p.fc2b = take(out);
  p.rrow = take(C2);
  p.hrow = take(Hd);
  p.dhrow = take(Hd);
  p.drrow = take(h2);
  p.prow = take(out);
  p.red = take((S2_THREADS / 32) * h2);
  p.arg0 = take(k8 * h1);
  p.arg1 = take(q8 * h2);
  p.rp0 = take(max_n + 1);
  p.col0 = take(max_e);
  p.rp1 = take(max_k + 1);
  p.col1 = take(p.max_e1);
  p.cmp0 = take(max_k + 1);
  p.cmem0 = take(max_n);
  p.cl0 = take(max_n);
  p.cmp1 = take(max_q + 1);
  p.cmem1 = take(max_k);
  p.cl1 = take(max_k);
  p.cscp1 = take(max_k + 1);
  p.cscr1 = take(p.max_e1);
  p.total = o;
  return p;
}

// C[m][n0..n0+3] = sum_k A[m*lda + k] * Bm[k*ldb + n]   (A row-major, K % 4 == 0, N % 4 == 0).
// TM x 4 register tile; threads [t0, t0+nth) of the CTA take part.  Each output element is one fmaf
// chain over ascending k.  out(m, n0, float4) is called for rows m < M only.
template <int TM, typename FO>
__device__ __forceinline__ void s2_gemm_rowA(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, int M,
                                             int N, int K, int tid, int nth, FO out) {
  const int mt = (M + TM - 1) / TM, nt = N >> 2;
  for (int item = tid; item < mt * nt; item += nth) {
    const int mg = item / nt, ng = item - mg * nt;
    float4 acc[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* ap = A + (mg * TM) * lda;
    const float* bp = Bm + ng * 4;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
      const float4 b0 = *reinterpret_cast<const float4*>(bp + (k + 0) * ldb);
      const float4 b1 = *reinterpret_cast<const float4*>(bp + (k + 1) * ldb);
      const float4 b2 = *reinterpret_cast<const float4*>(bp + (k + 2) * ldb);
      const float4 b3 = *reinterpret_cast<const float4*>(bp + (k + 3) * ldb);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float4 a = *reinterpret_cast<const float4*>(ap + i * lda + k);
        acc[i].x = fmaf(a.x, b0.x, acc[i].x); acc[i].y = fmaf(a.x, b0.y, acc[i].y);
        acc[i].z = fmaf(a.x, b0.z, acc[i].z); acc[i].w = fmaf(a.x, b0.w, acc[i].w);
        acc[i].x = fmaf(a.y, b1.x, acc[i].x); acc[i].y = fmaf(a.y, b1.y, acc[i].y);
        acc[i].z = fmaf(a.y, b1.z, acc[i].z); acc[i].w = fmaf(a.y, b1.w, acc[i].w);
        acc[i].x = fmaf(a.z, b2.x, acc[i].x); acc[i].y = fmaf(a.z, b2.y, acc[i].y);
        acc[i].z = fmaf(a.z, b2.z, acc[i].z); acc[i].w = fmaf(a.z, b2.w, acc[i].w);
        acc[i].x = fmaf(a.w, b3.x, acc[i].x); acc[i].y = fmaf(a.w, b3.y, acc[i].y);
        acc[i].z = fmaf(a.w, b3.z, acc[i].z); acc[i].w = fmaf(a.w, b3.w, acc[i].w);
      }
    }
#pragma unroll
    for (int i = 0; i < TM; ++i)
      if (mg * TM + i < M) out(mg * TM + i, ng * 4, acc[i]);
  }
}

// Split-K partial products of C[m][n] = sum_{k<K} At[k*lda + m] * Bm[k*ldb + n]  (M % 4 == 0, N % 4 == 0):
// split s in [0, KS) covers k in [s*chunk, min(K, (s+1)*chunk)) and stores its 4x4 tiles to
// scratch[s][M][N].  s2_splitk_reduce sums the KS slices in ascending s.
__device__ __forceinline__ void s2_splitk_partial(const float* __restrict__ At, int lda, const float* __restrict__ Bm, int ldb, int M,
                                                  int N, int K, int KS, float* __restrict__ scratch, int tid, int nth) {
  const int mt = M >> 2, nt = N >> 2, tiles = mt * nt;
  const int chunk = (K + KS - 1) / KS;
  for (int item = tid; item < tiles * KS; item += nth) {
    const int s = item / tiles, tile = item - s * tiles;
    const int mg = tile / nt, ng = tile - mg * nt;
    const int kb = s * chunk, ke = min(K, kb + chunk);
    float4 acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* ap = At + mg * 4;
    const float* bp = Bm + ng * 4;
#pragma unroll 4
    for (int k = kb; k < ke; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(ap + k * lda);
      const float4 b = *reinterpret_cast<const float4*>(bp + k * ldb);
      acc[0].x = fmaf(a.x, b.x, acc[0].x); acc[0].y = fmaf(a.x, b.y, acc[0].y);
      acc[0].z = fmaf(a.x, b.z, acc[0].z); acc[0].w = fmaf(a.x, b.w, acc[0].w);
      acc[1].x = fmaf(a.y, b.x, acc[1].x); acc[1].y = fmaf(a.y, b.y, acc[1].y);
      acc[1].z = fmaf(a.y, b.z, acc[1].z); acc[1].w = fmaf(a.y, b.w, acc[1].w);
      acc[2].x = fmaf(a.z, b.x, acc[2].x); acc[2].y = fmaf(a.z, b.y, acc[2].y);
      acc[2].z = fmaf(a.z, b.z, acc[2].z); acc[2].w = fmaf(a.z, b.w, acc[2].w);
      acc[3].x = fmaf(a.w, b.x, acc[3].x); acc[3].y = fmaf(a.w, b.y, acc[3].y);
      acc[3].z = fmaf(a.w, b.z, acc[3].z); acc[3].w = fmaf(a.w, b.w, acc[3].w);
    }
    float* sp = scratch + (size_t)s * M * N + (mg * 4) * N + ng * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i) *reinterpret_cast<float4*>(sp + i * N) = acc[i];
  }
}
__device__ __forceinline__ void s2_splitk_reduce(const float* __restrict__ scratch, int MN, int KS, float* __restrict__ dst, int tid,
                                                 int nth) {
  for (int e = tid; e < MN; e += nth) {
    float acc = 0.f;
    for (int s = 0; s < KS; ++s) acc += scratch[(size_t)s * MN + e];
    dst[e] = acc;
  }
}

__host__ __device__ inline int s2_split(int cap_words, int mn) {
  int ks = cap_words / mn;
  if (ks > S2_MAX_KS) ks = S2_MAX_KS;
  if (ks < 1) ks = 1;
  return ks;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(S2_THREADS, 1)
    ginet_graph_step2_kernel(const drgnn_ginet_step_args s) {
  extern __shared__ __align__(16) float sm[];
  cgx::cluster_group cluster = cgx::this_cluster();
  const drgnn_ginet_fused_args& a = s.g;
  const int r = (int)cluster.block_rank();   // branch of this CTA
  const int g = blockIdx.x >> 1;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  constexpr int T = S2_THREADS, NW = S2_THREADS / 32;
  const int F = a.F, H1 = a.h1, H2 = a.h2, C1 = 2 * H1, C2 = 2 * H2, Hd = s.Hd, out = s.out;
  const int co1 = r * H1, co2 = r * H2;
  const bool mirror = (s.flags & 1) != 0;   // also store the intermediates to global memory
  const Step2Plan P = step2_plan(F, H1, H2, a.max_n, a.max_k, a.max_q, s.max_e, Hd, out);
  DRGNN_PHASE(0);
  float* xs = sm + P.xs;   float* ax = sm + P.ax;   float* z1 = sm + P.z1;   float* dz1 = sm + P.dz1;
  float* w1t = sm + P.w1t; float* w2t = sm + P.w2t; float* w2 = sm + P.w2;
  float* p1 = sm + P.p1;   float* ap = sm + P.ap;   float* z2 = sm + P.z2;   float* p2 = sm + P.p2;
  float* dz2 = sm + P.dz2; float* dap = sm + P.dap; float* dp1 = sm + P.dp1;
  float* fc1w = sm + P.fc1w; float* fc2w = sm + P.fc2w; float* fc1b = sm + P.fc1b; float* fc2b = sm + P.fc2b;
  float* rrow = sm + P.rrow; float* hrow = sm + P.hrow; float* dhrow = sm + P.dhrow; float* drrow = sm + P.drrow;
  float* prow = sm + P.prow; float* red = sm + P.red;
  int* ism = reinterpret_cast<int*>(sm);
  int* arg0 = ism + P.arg0;  int* arg1 = ism + P.arg1;
  int* rp0 = ism + P.rp0;    int* col0 = ism + P.col0;   int* rp1 = ism + P.rp1;   int* col1 = ism + P.col1;
  int* cmp0 = ism + P.cmp0;  int* cmem0 = ism + P.cmem0; int* cl0 = ism + P.cl0;
  int* cmp1 = ism + P.cmp1;  int* cmem1 = ism + P.cmem1; int* cl1 = ism + P.cl1;
  int* cscp1 = ism + P.cscp1; int* cscr1 = ism + P.cscr1;
  const int LDX = P.ldx, LDZ1 = P.ldz1, LDP = P.ldp, LDZ2 = P.ldz2;

  // ---- graph extents (two dependent levels of tiny loads, the only global latency chain of the kernel)
  const int n0 = __ldg(a.node_ptr + g), n = __ldg(a.node_ptr + g + 1) - n0;
  const int k0 = __ldg(a.kptr0 + g), K = __ldg(a.kptr0 + g + 1) - k0;
  const int q0 = __ldg(a.kptr1 + g), Q = __ldg(a.kptr1 + g + 1) - q0;
  const bool train = !(s.forward_only || s.task == 0);
  float* part = s.partial + (int64_t)g * s.partial_ld;
  // loop-invariant global scalars of the head, fetched while the staging copies fly
  const uint32_t drop_ctr = (!s.keep && s.drop_p > 0.f) ? (uint32_t)__ldg(s.step_dev) : 0u;
  float y_first = 0.f;
  int y_cls = 0;
  if (train && t == 0) {
    if (s.task == 3) y_cls = (int)__ldg(s.y_class + g);
    else y_first = __ldg(s.y + (int64_t)g * out);
  }
  bool ok = n >= 0 && K >= 0 && Q >= 0 && n <= a.max_n && K <= a.max_k && Q <= a.max_q;
  int e00 = 0, E0 = 0, e10 = 0, E1 = 0, c10 = 0, m00 = 0, m10 = 0;
  if (ok) {
    // feature tile and head weights first: they need nothing but n0 / n
    {
      const float4* src = reinterpret_cast<const float4*>(a.x + (int64_t)n0 * F);
      float4* dst = reinterpret_cast<float4*>(xs);
      for (int i = t; i < n * (F >> 2); i += T) s2_cp16(dst + i, src + i);
    }
    e00 = __ldg(a.rowptr0 + n0); E0 = __ldg(a.rowptr0 + n0 + n) - e00;
    e10 = __ldg(a.rowptr1 + k0); E1 = __ldg(a.rowptr1 + k0 + K) - e10;
    c10 = __ldg(a.cscptr1 + k0);
    m00 = __ldg(a.cmptr0 + k0);  m10 = __ldg(a.cmptr1 + q0);
    ok = E0 >= 0 && E1 >= 0 && E0 <= s.max_e && E1 <= P.max_e1;
  }
  if (!ok) {   // host bounds violated: flag and leave (checked by validate()); both CTAs take this branch
    s2_wait<0>();
    if (t == 0) atomicOr(a.status, 64);
    if (train && r == 0)
      for (int i = t; i < s.n_params + 1; i += T) part[i] = 0.f;
    return;
  }
  // The peer CTA must be running before its shared memory is written (read-out exchange below):
  // arrive now, wait just before the exchange, so the barrier costs nothing.
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  // ---- group 0 (forward): index slices of both levels
  for (int i = t; i <= n; i += T) s2_cp4(rp0 + i, a.rowptr0 + n0 + i);
  for (int i = t; i < E0; i += T) s2_cp4(col0 + i, a.col0 + e00 + i);
  for (int i = t; i <= K; i += T) s2_cp4(rp1 + i, a.rowptr1 + k0 + i);
  for (int i = t; i < E1; i += T) s2_cp4(col1 + i, a.col1 + e10 + i);
  for (int i = t; i <= K; i += T) s2_cp4(cmp0 + i, a.cmptr0 + k0 + i);
  for (int i = t; i < n; i += T) s2_cp4(cmem0 + i, a.cmem0 + m00 + i);
  for (int i = t; i <= Q; i += T) s2_cp4(cmp1 + i, a.cmptr1 + q0 + i);
  for (int i = t; i < K; i += T) s2_cp4(cmem1 + i, a.cmem1 + m10 + i);
  s2_commit();
  // ---- group 1 (head + backward): head weights, cluster ids, CSC of the coarsened graph
  {
    const float4* src = reinterpret_cast<const float4*>(s.fc1_w);
    float4* dst = reinterpret_cast<float4*>(fc1w);
    for (int i = t; i < (Hd * C2) >> 2; i += T) s2_cp16(dst + i, src + i);
    for (int i = t; i < out * Hd; i += T) s2_cp4(fc2w + i, s.fc2_w + i);
    if (s.fc1_b)
      for (int i = t; i < Hd; i += T) s2_cp4(fc1b + i, s.fc1_b + i);
    if (s.fc2_b)
      for (int i = t; i < out; i += T) s2_cp4(fc2b + i, s.fc2_b + i);
    if (train) {
      for (int i = t; i < n; i += T) s2_cp4(cl0 + i, a.cl0 + n0 + i);
      for (int i = t; i < K; i += T) s2_cp4(cl1 + i, a.cl1 + k0 + i);
      for (int i = t; i <= K; i += T) s2_cp4(cscp1 + i, a.cscptr1 + k0 + i);
      for (int i = t; i < E1; i += T) s2_cp4(cscr1 + i, a.cscrow1 + c10 + i);
    }
  }
  s2_commit();
  // ---- this branch's weights, transposed through registers
  for (int i = t; i < H1 * F; i += T) {        // W1 [C1][F] rows co1.. -> w1t [F][H1]
    const int c = i / F, f = i - c * F;
    w1t[f * H1 + c] = __ldg(a.W1 + (int64_t)(co1 + c) * F + f);
  }
  for (int i = t; i < H2 * H1; i += T) {       // W2 [2][H2][H1] group r -> w2 [H2][H1], w2t [H1][H2]
    const int o = i / H1, j = i - o * H1;
    const float v = __ldg(a.W2 + (int64_t)r * H2 * H1 + i);
    w2[i] = v;
    w2t[j * H2 + o] = v;
  }
  if (!s.fc1_b)
    for (int i = t; i < Hd; i += T) fc1b[i] = 0.f;
  if (!s.fc2_b)
    for (int i = t; i < out; i += T) fc2b[i] = 0.f;
  s2_wait<1>();
  __syncthreads();
  DRGNN_PHASE(1);

  const int F4 = F >> 2, H14 = H1 >> 2, H24 = H2 >> 2;
  // ---- AX = A x : the F/4 lanes of a row share its edge list and read whole 16-byte-aligned feature rows
  for (int item = t; item < n * F4; item += T) {
    const int i = item / F4, q4 = item - i * F4;
    const int sb = rp0[i] - e00, se = rp0[i + 1] - e00;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int p = sb; p < se; ++p) {
      const int c = col0[p] - n0;
      const float4 v = *reinterpret_cast<const float4*>(xs + c * F + q4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(ax + i * LDX + q4 * 4) = acc;
    if (mirror && r == 0) *reinterpret_cast<float4*>(a.Zin1 + (int64_t)(n0 + i) * F + q4 * 4) = acc;
  }
  __syncthreads();
  DRGNN_PHASE(2);
  // ---- Z1 = relu(AX W1_r^T)
  s2_gemm_rowA<2>(ax, LDX, w1t, H1, n, H1, F, t, T, [&](int m, int c, float4 v) {
    v.x = v.x < 0.f ? 0.f : v.x; v.y = v.y < 0.f ? 0.f : v.y; v.z = v.z < 0.f ? 0.f : v.z; v.w = v.w < 0.f ? 0.f : v.w;
    *reinterpret_cast<float4*>(z1 + m * LDZ1 + c) = v;
    if (mirror) *reinterpret_cast<float4*>(a.Z1 + (int64_t)(n0 + m) * C1 + co1 + c) = v;
  });
  __syncthreads();
  DRGNN_PHASE(3);
  // ---- P1 = cluster max of Z1 (first member wins ties, a NaN never wins; community_pooling.py:201)
  for (int item = t; item < K * H14; item += T) {
    const int k = item / H14, q4 = item - k * H14;
    const int sb = cmp0[k] - m00, se = cmp0[k + 1] - m00;
    float4 best = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    int4 arg = make_int4(-1, -1, -1, -1);
    for (int p = sb; p < se; ++p) {
      const int i = cmem0[p];
      const float4 v = *reinterpret_cast<const float4*>(z1 + (i - n0) * LDZ1 + q4 * 4);
      if (v.x > best.x) { best.x = v.x; arg.x = i; }
      if (v.y > best.y) { best.y = v.y; arg.y = i; }
      if (v.z > best.z) { best.z = v.z; arg.z = i; }
      if (v.w > best.w) { best.w = v.w; arg.w = i; }
    }
    if (arg.x < 0) best.x = 0.f;
    if (arg.y < 0) best.y = 0.f;
    if (arg.z < 0) best.z = 0.f;
    if (arg.w < 0) best.w = 0.f;
    *reinterpret_cast<float4*>(p1 + k * LDP + q4 * 4) = best;
    *reinterpret_cast<int4*>(arg0 + k * H1 + q4 * 4) = arg;
    if (mirror) *reinterpret_cast<int4*>(a.arg0 + (int64_t)(k0 + k) * C1 + co1 + q4 * 4) = arg;
  }
  __syncthreads();
  DRGNN_PHASE(4);
  // ---- AP = A1 P1 on the coarsened graph
  for (int item = t; item < K * H14; item += T) {
    const int k = item / H14, q4 = item - k * H14;
    const int sb = rp1[k] - e10, se = rp1[k + 1] - e10;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int p = sb; p < se; ++p) {
      const int c = col1[p] - k0;
      const float4 v = *reinterpret_cast<const float4*>(p1 + c * LDP + q4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(ap + k * LDP + q4 * 4) = acc;
    if (mirror) *reinterpret_cast<float4*>(a.Zin2 + (int64_t)(k0 + k) * C1 + co1 + q4 * 4) = acc;
  }
  __syncthreads();
  DRGNN_PHASE(5);
  // ---- Z2 = relu(AP W2_r^T)
  s2_gemm_rowA<2>(ap, LDP, w2t, H2, K, H2, H1, t, T, [&](int m, int o, float4 v) {
    v.x = v.x < 0.f ? 0.f : v.x; v.y = v.y < 0.f ? 0.f : v.y; v.z = v.z < 0.f ? 0.f : v.z; v.w = v.w < 0.f ? 0.f : v.w;
    *reinterpret_cast<float4*>(z2 + m * LDZ2 + o) = v;
    if (mirror) *reinterpret_cast<float4*>(a.Z2 + (int64_t)(k0 + m) * C2 + co2 + o) = v;
  });
  __syncthreads();
  DRGNN_PHASE(6);
  // ---- P2 = level-1 cluster max (max_pool_x)
  for (int item = t; item < Q * H24; item += T) {
    const int q = item / H24, q4 = item - q * H24;
    const int sb = cmp1[q] - m10, se = cmp1[q + 1] - m10;
    float4 best = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    int4 arg = make_int4(-1, -1, -1, -1);
    for (int p = sb; p < se; ++p) {
      const int k = cmem1[p];
      const float4 v = *reinterpret_cast<const float4*>(z2 + (k - k0) * LDZ2 + q4 * 4);
      if (v.x > best.x) { best.x = v.x; arg.x = k; }
      if (v.y > best.y) { best.y = v.y; arg.y = k; }
      if (v.z > best.z) { best.z = v.z; arg.z = k; }
      if (v.w > best.w) { best.w = v.w; arg.w = k; }
    }
    if (arg.x < 0) best.x = 0.f;
    if (arg.y < 0) best.y = 0.f;
    if (arg.z < 0) best.z = 0.f;
    if (arg.w < 0) best.w = 0.f;
    *reinterpret_cast<float4*>(p2 + q * H2 + q4 * 4) = best;
    *reinterpret_cast<int4*>(arg1 + q * H2 + q4 * 4) = arg;
    if (mirror) *reinterpret_cast<int4*>(a.arg1 + (int64_t)(q0 + q) * C2 + co2 + q4 * 4) = arg;
  }
  __syncthreads();
  DRGNN_PHASE(7);
  // ---- R[g] half = mean over the graph's level-1 clusters; both halves land in both CTAs (DSMEM)
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  {
    float* peer_rrow = cluster.map_shared_rank(rrow, (unsigned)(r ^ 1));
    for (int c = t; c < H2; c += T) {
      float acc = 0.f;
      for (int q = 0; q < Q; ++q) acc += p2[q * H2 + c];
      acc *= 1.f / (float)max(Q, 1);
      a.R[(int64_t)g * C2 + co2 + c] = acc;
      rrow[co2 + c] = acc;
      peer_rrow[co2 + c] = acc;
    }
  }
  s2_wait<0>();      // head weights / backward indices have landed (this thread's copies)
  cluster.sync();    // ... everybody's, and the peer's half of the read-out row
  DRGNN_PHASE(8);
  // ---- fc1: warp per hidden unit, lanes over the read-out channels (both CTAs, identical results)
  for (int j = warp; j < Hd; j += NW) {
    float acc = 0.f;
    for (int c = lane; c < C2; c += 32) acc = fmaf(rrow[c], fc1w[j * C2 + c], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += fc1b[j];
      acc = acc < 0.f ? 0.f : acc;
      if (s.keep) {
        acc = s.keep[(int64_t)g * Hd + j] > 0.f ? acc * s.keep_scale : 0.f;
      } else if (s.drop_p > 0.f) {
        acc = hash_uniform(s.seed, drop_ctr, (uint32_t)(g * Hd + j)) >= s.drop_p ? acc * s.keep_scale : 0.f;
      }
      hrow[j] = acc;
    }
  }
  __syncthreads();
  DRGNN_PHASE(9);
  // ---- fc2: warp per output
  for (int o = warp; o < out; o += NW) {
    float acc = 0.f;
    for (int j = lane; j < Hd; j += 32) acc = fmaf(hrow[j], fc2w[o * Hd + j], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += fc2b[o];
      prow[o] = acc;
      if (r == 0) s.pred[(int64_t)g * out + o] = acc;
    }
  }
  __syncthreads();
  DRGNN_PHASE(10);
  if (!train) return;
  // ---- loss term of this graph and dLoss/dpred (one thread per CTA, identical results)
  if (t == 0) {
    float lg = 0.f;
    if (s.task == 3) {
      float mx = prow[0];
      for (int c = 1; c < out; ++c) mx = fmaxf(mx, prow[c]);
      float se = 0.f;
      for (int c = 0; c < out; ++c) se += expf(prow[c] - mx);
      const float lse = mx + logf(se);
      const int tc = y_cls;
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
        const float d = p - (c == 0 ? y_first : s.y[(int64_t)g * out + c]);
        lg += d * d;
        prow[c] = 2.f * d * s.inv_norm * dp;
      }
    }
    if (r == 0) part[s.n_params] = lg * s.inv_norm;   // summed into the loss by the reduce kernel
  }
  __syncthreads();
  // ---- head backward: dh (all units, both CTAs); gradient rows of the hidden units [j0, j1) by CTA r
  const int j0 = r ? (Hd >> 1) : 0, j1 = r ? Hd : (Hd >> 1), nj = j1 - j0;
  for (int j = t; j < Hd; j += T) {
    float acc = 0.f;
    for (int o = 0; o < out; ++o) acc = fmaf(prow[o], fc2w[o * Hd + j], acc);
    acc = hrow[j] > 0.f ? acc * s.keep_scale : 0.f;
    dhrow[j] = acc;
    if (j >= j0 && j < j1) part[s.off_fc1b + j] = acc;
  }
  for (int i = t; i < out * nj; i += T) {
    const int o = i / nj, j = j0 + (i - o * nj);
    part[s.off_fc2w + o * Hd + j] = prow[o] * hrow[j];
  }
  if (r == 0)
    for (int o = t; o < out; o += T) part[s.off_fc2b + o] = prow[o];
  __syncthreads();
  for (int i = t; i < nj * C2; i += T) {
    const int j = j0 + i / C2, c = i % C2;
    part[s.off_fc1w + j * C2 + c] = dhrow[j] * rrow[c];
  }
  // dR[c] of this branch's channels = sum_j dh[j] W1[j][co2 + c]: warps split the hidden units, fixed-order sum
  for (int c = lane; c < H2; c += 32) {
    float acc = 0.f;
    for (int j = warp; j < Hd; j += NW) acc = fmaf(dhrow[j], fc1w[j * C2 + co2 + c], acc);
    red[warp * H2 + c] = acc;
  }
  __syncthreads();
  for (int c = t; c < H2; c += T) {
    float acc = 0.f;
    for (int w = 0; w < NW; ++w) acc += red[w * H2 + c];
    drrow[c] = acc;
  }
  __syncthreads();
  DRGNN_PHASE(11);
  // ---- dZ2: read-out mean backward, routed to the arg-max member, gated by ReLU
  {
    const float invQ = 1.f / (float)max(Q, 1);
    for (int item = t; item < K * H24; item += T) {
      const int k = item / H24, q4 = item - k * H24;
      const int q = cl1[k] - q0;
      const int4 am = *reinterpret_cast<const int4*>(arg1 + q * H2 + q4 * 4);
      const float4 z = *reinterpret_cast<const float4*>(z2 + k * LDZ2 + q4 * 4);
      const float4 d = *reinterpret_cast<const float4*>(drrow + q4 * 4);
      const int me = k0 + k;
      float4 v;
      v.x = (am.x == me && z.x > 0.f) ? d.x * invQ : 0.f;
      v.y = (am.y == me && z.y > 0.f) ? d.y * invQ : 0.f;
      v.z = (am.z == me && z.z > 0.f) ? d.z * invQ : 0.f;
      v.w = (am.w == me && z.w > 0.f) ? d.w * invQ : 0.f;
      *reinterpret_cast<float4*>(dz2 + k * LDZ2 + q4 * 4) = v;
    }
  }
  __syncthreads();
  DRGNN_PHASE(12);
  // ---- dW2_r = dZ2^T AP (split over the K0 rows, first half of the CTA)  ||  dAP = dZ2 W2_r (second half)
  const int KS2 = s2_split(P.xs_words, H2 * H1), KS1 = s2_split(P.xs_words, H1 * F);
  float* scratch = xs;   // the feature tile is dead since AX
  if (t < (T >> 1)) {
    s2_splitk_partial(dz2, LDZ2, ap, LDP, H2, H1, K, KS2, scratch, t, T >> 1);
  } else {
    s2_gemm_rowA<2>(dz2, LDZ2, w2, H1, K, H1, H2, t - (T >> 1), T >> 1,
                    [&](int m, int j, float4 v) { *reinterpret_cast<float4*>(dap + m * LDP + j) = v; });
  }
  __syncthreads();
  DRGNN_PHASE(13);
  s2_splitk_reduce(scratch, H2 * H1, KS2, part + s.off_w2 + r * H2 * H1, t, T);
  // ---- dP1 = A1^T dAP  (CSC of the coarsened graph)
  for (int item = t; item < K * H14; item += T) {
    const int k = item / H14, q4 = item - k * H14;
    const int sb = cscp1[k] - c10, se = cscp1[k + 1] - c10;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
    for (int p = sb; p < se; ++p) {
      const int rr = cscr1[p] - k0;
      const float4 v = *reinterpret_cast<const float4*>(dap + rr * LDP + q4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(dp1 + k * LDP + q4 * 4) = acc;
  }
  __syncthreads();
  DRGNN_PHASE(14);
  // ---- dZ1: routed to the arg-max node of its cluster, gated by ReLU
  for (int item = t; item < n * H14; item += T) {
    const int i = item / H14, q4 = item - i * H14;
    const int k = cl0[i] - k0;
    const int4 am = *reinterpret_cast<const int4*>(arg0 + k * H1 + q4 * 4);
    const float4 z = *reinterpret_cast<const float4*>(z1 + i * LDZ1 + q4 * 4);
    const float4 d = *reinterpret_cast<const float4*>(dp1 + k * LDP + q4 * 4);
    const int me = n0 + i;
    float4 v;
    v.x = (am.x == me && z.x > 0.f) ? d.x : 0.f;
    v.y = (am.y == me && z.y > 0.f) ? d.y : 0.f;
    v.z = (am.z == me && z.z > 0.f) ? d.z : 0.f;
    v.w = (am.w == me && z.w > 0.f) ? d.w : 0.f;
    *reinterpret_cast<float4*>(dz1 + i * LDZ1 + q4 * 4) = v;
  }
  __syncthreads();
  DRGNN_PHASE(15);
  // ---- dW1_r [H1][F] = dZ1^T AX, split over the nodes
  s2_splitk_partial(dz1, LDZ1, ax, LDX, H1, F, n, KS1, scratch, t, T);
  __syncthreads();
  s2_splitk_reduce(scratch, H1 * F, KS1, part + s.off_w1 + r * H1 * F, t, T);
  DRGNN_PHASE(16);
}

static inline bool step2_shapes_ok(const drgnn_ginet_step_args& s) {
  const drgnn_ginet_fused_args& a = s.g;
  return a.nb == 2 && a.F % 4 == 0 && a.h1 % 4 == 0 && a.h2 % 4 == 0 && s.max_e > 0 && a.max_n > 0 && a.max_k > 0 &&
         a.max_q > 0 && s.Hd > 0 && s.out > 0 && ((2 * a.h2 * s.Hd) % 4 == 0);
}
