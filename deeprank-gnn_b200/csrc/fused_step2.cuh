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
//   * ONE asynchronous staging pass - three TMA bulk copies (cp.async.bulk + mbarrier): the graph's
//     structure blob (EVERY index slice of both graph levels with graph-local ids, written by the
//     structure pass, structure_blob.cu), its feature tile and fc1.weight - brings everything into
//     shared memory; after it no phase touches global memory except to store results;
//   * every intermediate the backward needs (AX, Z1, argmax0, AP, Z2, argmax1) stays in shared
//     memory (optionally mirrored to global memory for the parity tests, flag bit 0);
//   * the two halves of the read-out row are exchanged through distributed shared memory
//     (one cluster barrier); the tiny head (fc1 / ReLU / dropout / fc2 / loss) is evaluated by both
//     CTAs, its gradient rows are written half by each;
//   * the weight-gradient products (dW1 = dZ1^T AX, dW2 = dZ2^T AP: M, N tiny, K = nodes) are
//     split over K across the whole CTA and reduced in a fixed order (deterministic);
//   * when the grid is co-resident (B <= clusters the device holds) the per-graph gradient rows are
//     reduced behind a grid barrier inside the same launch, Adam is applied there, and on several
//     GPUs the slices travel to the peers over NVLink as 8-byte {value, epoch} words first
//     (rank-ordered sum: bit-identical weights on every rank, no flag, no fence, no extra launch).
//
// Arithmetic per output element of the graph part is the same fmaf chain, in the same order, as v1 /
// the op-level kernels; the split-K weight gradients, the 4-lane fc1 and the 4-quarter gradient
// reduction use their own fixed orders.
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
__device__ __forceinline__ unsigned s2_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void s2_st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t s2_ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// low-latency exchange word: value bits and epoch in ONE 8-byte store / load (never torn)
__device__ __forceinline__ void s2_st_ll(uint64_t* p, unsigned bits, unsigned epoch) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(bits), "r"(epoch) : "memory");
}
__device__ __forceinline__ void s2_ld_ll(const uint64_t* p, unsigned& bits, unsigned& epoch) {
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(bits), "=r"(epoch) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long s2_globaltimer() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}
// TMA bulk copy global -> shared, completion on an mbarrier (bytes % 16 == 0, both sides 16-byte aligned)
__device__ __forceinline__ void s2_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s2_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void s2_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s2_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s2_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s2_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(s2_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void s2_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "S2_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra S2_DONE_%=;\n"
      "bra S2_WAIT_%=;\n"
      "S2_DONE_%=:\n"
      "}\n" ::"r"(s2_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
template <int N>
__device__ __forceinline__ void s2_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__host__ __device__ inline int s2_up4(int x) { return (x + 3) & ~3; }
// item -> (item / w, item % w) with a shift when w is a power of two (the usual widths: 4, 8)
__host__ __device__ inline int s2_log2_exact(int w) {
  int sh = 0;
  while ((1 << sh) < w) ++sh;
  return (1 << sh) == w ? sh : -1;
}
struct S2Div {
  int w, sh;
  __device__ __forceinline__ explicit S2Div(int w_) : w(w_), sh(s2_log2_exact(w_)) {}
  __device__ __forceinline__ void split(int item, int& q, int& r) const {
    if (sh >= 0) { q = item >> sh; r = item & (w - 1); }
    else { q = item / w; r = item - q * w; }
  }
};

// Shared-memory plan (offsets in 4-byte words, every offset a multiple of 4 = 16 bytes).
struct Step2Plan {
  // float regions
  int xs, ax, z1, dz1, w1t, w2t, w2, p1, ap, z2, p2, dz2, dap, dp1;
  int fc1w, fc2w, fc1b, fc2b, rrow, hrow, dhrow, drrow, prow, red;
  // int regions: arg-max ids, the graph's structure blob (every index slice, staged by one bulk copy), mbarriers
  int arg0, arg1, blob, bars;
  int blob_words;
  int ldx, ldz1, ldp, ldz2;   // row strides (words): F+4, h1+4, h1+4, h2+4
  int xs_words;               // capacity of the xs region (split-K scratch in the backward)
  int max_e1;
  int total;
  int fused_reduce;           // set by the launcher: the gradient reduction (+ Adam) runs inside the launch
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
  p.blob_words = DRGNN_BLOB_USED(max_n, max_e);
  p.blob = take(p.blob_words);
  p.bars = take(4);           // two 8-byte mbarriers
  p.total = o;
  p.fused_reduce = 0;
  return p;
}

// The phases below are __noinline__ on purpose: the kernel runs every phase ONCE per CTA, so
// straight-line inlined code is paid for in instruction-cache misses (the inlined version of this
// kernel was 9.5 k SASS instructions = 150 KB, about 5 cycles per instruction, every phase fetch
// bound).  Shared routines keep the code small and warm: the second and third use run from cache.

// C[m][n..n+3] = (relu) sum_k A[m*lda + k] * Bm[k*ldb + n]   (A row-major, K % 4 == 0, N % 4 == 0).
// 2 x 4 register tile; threads tid in [0, nth) take part.  Each output element is one fmaf chain
// over ascending k (bit-identical to the op-level linear kernel).  Rows m >= M are not written.
__device__ __noinline__ void s2_gemm(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, int M, int N, int K,
                                     float* __restrict__ C, int ldc, int relu, int tid, int nth) {
  const int mt = (M + 1) >> 1, nt = N >> 2;
#pragma unroll 1
  for (int item = tid; item < mt * nt; item += nth) {
    const int mg = item / nt, ng = item - mg * nt;
    float4 acc0 = make_float4(0.f, 0.f, 0.f, 0.f), acc1 = acc0;
    const float* ap = A + (mg * 2) * lda;
    const float* bp = Bm + ng * 4;
#pragma unroll 1
    for (int k = 0; k < K; k += 4) {
      const float4 b0 = *reinterpret_cast<const float4*>(bp + (k + 0) * ldb);
      const float4 b1 = *reinterpret_cast<const float4*>(bp + (k + 1) * ldb);
      const float4 b2 = *reinterpret_cast<const float4*>(bp + (k + 2) * ldb);
      const float4 b3 = *reinterpret_cast<const float4*>(bp + (k + 3) * ldb);
      const float4 a0 = *reinterpret_cast<const float4*>(ap + k);
      const float4 a1 = *reinterpret_cast<const float4*>(ap + lda + k);
      acc0.x = fmaf(a0.x, b0.x, acc0.x); acc0.y = fmaf(a0.x, b0.y, acc0.y); acc0.z = fmaf(a0.x, b0.z, acc0.z); acc0.w = fmaf(a0.x, b0.w, acc0.w);
      acc1.x = fmaf(a1.x, b0.x, acc1.x); acc1.y = fmaf(a1.x, b0.y, acc1.y); acc1.z = fmaf(a1.x, b0.z, acc1.z); acc1.w = fmaf(a1.x, b0.w, acc1.w);
      acc0.x = fmaf(a0.y, b1.x, acc0.x); acc0.y = fmaf(a0.y, b1.y, acc0.y); acc0.z = fmaf(a0.y, b1.z, acc0.z); acc0.w = fmaf(a0.y, b1.w, acc0.w);
      acc1.x = fmaf(a1.y, b1.x, acc1.x); acc1.y = fmaf(a1.y, b1.y, acc1.y); acc1.z = fmaf(a1.y, b1.z, acc1.z); acc1.w = fmaf(a1.y, b1.w, acc1.w);
      acc0.x = fmaf(a0.z, b2.x, acc0.x); acc0.y = fmaf(a0.z, b2.y, acc0.y); acc0.z = fmaf(a0.z, b2.z, acc0.z); acc0.w = fmaf(a0.z, b2.w, acc0.w);
      acc1.x = fmaf(a1.z, b2.x, acc1.x); acc1.y = fmaf(a1.z, b2.y, acc1.y); acc1.z = fmaf(a1.z, b2.z, acc1.z); acc1.w = fmaf(a1.z, b2.w, acc1.w);
      acc0.x = fmaf(a0.w, b3.x, acc0.x); acc0.y = fmaf(a0.w, b3.y, acc0.y); acc0.z = fmaf(a0.w, b3.z, acc0.z); acc0.w = fmaf(a0.w, b3.w, acc0.w);
      acc1.x = fmaf(a1.w, b3.x, acc1.x); acc1.y = fmaf(a1.w, b3.y, acc1.y); acc1.z = fmaf(a1.w, b3.z, acc1.z); acc1.w = fmaf(a1.w, b3.w, acc1.w);
    }
    if (relu) {
      acc0.x = acc0.x < 0.f ? 0.f : acc0.x; acc0.y = acc0.y < 0.f ? 0.f : acc0.y; acc0.z = acc0.z < 0.f ? 0.f : acc0.z; acc0.w = acc0.w < 0.f ? 0.f : acc0.w;
      acc1.x = acc1.x < 0.f ? 0.f : acc1.x; acc1.y = acc1.y < 0.f ? 0.f : acc1.y; acc1.z = acc1.z < 0.f ? 0.f : acc1.z; acc1.w = acc1.w < 0.f ? 0.f : acc1.w;
    }
    const int m = mg * 2;
    *reinterpret_cast<float4*>(C + m * ldc + ng * 4) = acc0;
    if (m + 1 < M) *reinterpret_cast<float4*>(C + (m + 1) * ldc + ng * 4) = acc1;
  }
}

// dst[i][4q..4q+3] = sum over the CSR entries p of row i of src[col[p] - nbase][4q..4q+3]  (ascending p:
// the CPU scatter order).  The W4 lanes of a row share its index loads; rowptr / col hold global ids.
__device__ __noinline__ void s2_gather(const int* __restrict__ rp, const int* __restrict__ col, int ebase, int nbase,
                                       const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd, int rows, int W4,
                                       int tid, int nth) {
  const S2Div dv(W4);
#pragma unroll 1
  for (int item = tid; item < rows * W4; item += nth) {
    int i, q4;
    dv.split(item, i, q4);
    const int sb = rp[i] - ebase, se = rp[i + 1] - ebase;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 2
    for (int p = sb; p < se; ++p) {
      const float4 v = *reinterpret_cast<const float4*>(src + (col[p] - nbase) * lds + q4 * 4);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    *reinterpret_cast<float4*>(dst + i * ldd + q4 * 4) = acc;
  }
}

// Cluster max + argmax (global member id): first member wins ties, a NaN never wins, empty -> 0
// (torch_scatter's CPU scatter_max; community_pooling.py:201, max_pool_x).
__device__ __noinline__ void s2_cluster_max(const int* __restrict__ cmp, const int* __restrict__ cmem, int mbase, int nbase,
                                            const float* __restrict__ src, int lds, float* __restrict__ dst, int ldd,
                                            int* __restrict__ arg, int ldarg, int rows, int W4, int tid, int nth) {
  const S2Div dv(W4);
#pragma unroll 1
  for (int item = tid; item < rows * W4; item += nth) {
    int k, q4;
    dv.split(item, k, q4);
    const int sb = cmp[k] - mbase, se = cmp[k + 1] - mbase;
    float4 best = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    int4 am = make_int4(-1, -1, -1, -1);
#pragma unroll 1
    for (int p = sb; p < se; ++p) {
      const int i = cmem[p];
      const float4 v = *reinterpret_cast<const float4*>(src + (i - nbase) * lds + q4 * 4);
      if (v.x > best.x) { best.x = v.x; am.x = i; }
      if (v.y > best.y) { best.y = v.y; am.y = i; }
      if (v.z > best.z) { best.z = v.z; am.z = i; }
      if (v.w > best.w) { best.w = v.w; am.w = i; }
    }
    if (am.x < 0) best.x = 0.f;
    if (am.y < 0) best.y = 0.f;
    if (am.z < 0) best.z = 0.f;
    if (am.w < 0) best.w = 0.f;
    *reinterpret_cast<float4*>(dst + k * ldd + q4 * 4) = best;
    *reinterpret_cast<int4*>(arg + k * ldarg + q4 * 4) = am;
  }
}

// Backward of cluster max + ReLU: dst[i][c] = (arg[cl[i] - kbase][c] == self0 + i && z[i][c] > 0) ?
// d[(cl[i] - kbase) * ldd + c] * scale : 0.   ldd == 0: one gradient row for every cluster (read-out mean).
__device__ __noinline__ void s2_route(const int* __restrict__ cl, int kbase, const int* __restrict__ arg, int ldarg,
                                      const float* __restrict__ z, int ldz, const float* __restrict__ d, int ldd, float scale,
                                      float* __restrict__ dst, int lddst, int rows, int W4, int self0, int tid, int nth) {
  const S2Div dv(W4);
#pragma unroll 1
  for (int item = tid; item < rows * W4; item += nth) {
    int i, q4;
    dv.split(item, i, q4);
    const int k = cl[i] - kbase;
    const int4 am = *reinterpret_cast<const int4*>(arg + k * ldarg + q4 * 4);
    const float4 zz = *reinterpret_cast<const float4*>(z + i * ldz + q4 * 4);
    const float4 dd = *reinterpret_cast<const float4*>(d + k * ldd + q4 * 4);
    const int me = self0 + i;
    float4 v;
    v.x = (am.x == me && zz.x > 0.f) ? dd.x * scale : 0.f;
    v.y = (am.y == me && zz.y > 0.f) ? dd.y * scale : 0.f;
    v.z = (am.z == me && zz.z > 0.f) ? dd.z * scale : 0.f;
    v.w = (am.w == me && zz.w > 0.f) ? dd.w * scale : 0.f;
    *reinterpret_cast<float4*>(dst + i * lddst + q4 * 4) = v;
  }
}

// Split-K partial products of C[m][n] = sum_{k<K} At[k*lda + m] * Bm[k*ldb + n]  (M % 4 == 0, N % 4 == 0):
// split s in [0, KS) covers k in [s*chunk, min(K, (s+1)*chunk)) and stores its 4x4 tiles to
// scratch[s][M][N].  s2_splitk_reduce sums the KS slices in ascending s.
__device__ __noinline__ void s2_splitk_partial(const float* __restrict__ At, int lda, const float* __restrict__ Bm, int ldb, int M,
                                               int N, int K, int KS, float* __restrict__ scratch, int tid, int nth) {
  const int mt = M >> 2, nt = N >> 2, tiles = mt * nt;
  const int chunk = (K + KS - 1) / KS;
#pragma unroll 1
  for (int item = tid; item < tiles * KS; item += nth) {
    const int s = item / tiles, tile = item - s * tiles;
    const int mg = tile / nt, ng = tile - mg * nt;
    const int kb = s * chunk, ke = min(K, kb + chunk);
    float4 acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* ap = At + mg * 4;
    const float* bp = Bm + ng * 4;
#pragma unroll 2
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
__device__ __noinline__ void s2_splitk_reduce(const float* __restrict__ scratch, int MN, int KS, float* __restrict__ dst, int tid,
                                              int nth) {
#pragma unroll 1
  for (int e = tid; e < MN; e += nth) {
    float acc = 0.f;
#pragma unroll 4
    for (int s = 0; s < KS; ++s) acc += scratch[(size_t)s * MN + e];
    dst[e] = acc;
  }
}

// asynchronous copy of `count` 4-byte words global -> shared (any alignment)
__device__ __noinline__ void s2_stage32(void* dst, const void* src, int count, int tid, int nth) {
  const uint32_t d = s2_smem_u32(dst);
  const char* g = reinterpret_cast<const char*>(src);
#pragma unroll 1
  for (int i = tid; i < count; i += nth)
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d + 4u * i), "l"(g + 4 * (size_t)i) : "memory");
}
// ... of `count` 16-byte words (both sides 16-byte aligned)
__device__ __noinline__ void s2_stage128(void* dst, const void* src, int count, int tid, int nth) {
  const uint32_t d = s2_smem_u32(dst);
  const char* g = reinterpret_cast<const char*>(src);
#pragma unroll 1
  for (int i = tid; i < count; i += nth)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * i), "l"(g + 16 * (size_t)i) : "memory");
}
// rows x W4 16-byte words of shared memory -> global memory (test mirror of the intermediates).
// iadd1 > 0: the words are int32 local ids; non-negative ones are stored as id + (iadd1 - 1).
__device__ __noinline__ void s2_mirror(const void* src, int lds, void* dst, int64_t ldd, int rows, int W4, int iadd1, int tid,
                                       int nth) {
  const int* sp = reinterpret_cast<const int*>(src);
  int* dp = reinterpret_cast<int*>(dst);
  const S2Div dv(W4);
#pragma unroll 1
  for (int item = tid; item < rows * W4; item += nth) {
    int i, q4;
    dv.split(item, i, q4);
    int4 v = *reinterpret_cast<const int4*>(sp + i * lds + q4 * 4);
    if (iadd1 > 0) {
      const int ad = iadd1 - 1;
      v.x = v.x >= 0 ? v.x + ad : v.x; v.y = v.y >= 0 ? v.y + ad : v.y;
      v.z = v.z >= 0 ? v.z + ad : v.z; v.w = v.w >= 0 ? v.w + ad : v.w;
    }
    *reinterpret_cast<int4*>(dp + i * ldd + q4 * 4) = v;
  }
}

__host__ __device__ inline int s2_split(int cap_words, int mn) {
  int ks = cap_words / mn;
  if (ks > S2_MAX_KS) ks = S2_MAX_KS;
  if (ks < 1) ks = 1;
  return ks;
}

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(S2_THREADS, 1)
    ginet_graph_step2_kernel(const drgnn_ginet_step_args s, const Step2Plan P, const drgnn_peer_comm C) {
  extern __shared__ __align__(16) float sm[];
  cgx::cluster_group cluster = cgx::this_cluster();
  const drgnn_ginet_fused_args& a = s.g;
  const int r = (int)cluster.block_rank();   // branch of this CTA
  const int g = blockIdx.x >> 1;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // phase clocks of block 0 (diagnostic, flag bit 3): their global stores delay the release fences of block 0, and
  // the whole grid waits for block 0 at the grid barrier, so they are off unless asked for
  const bool timers = (s.flags & 8) != 0 && blockIdx.x == 0 && t == 0;
#define S2_PHASE(i)                                                  \
  do {                                                               \
    if (timers) g_phase[i] = (unsigned long long)clock64();          \
  } while (0)
  constexpr int T = S2_THREADS, NW = S2_THREADS / 32;
  const int F = a.F, H1 = a.h1, H2 = a.h2, C1 = 2 * H1, C2 = 2 * H2, Hd = s.Hd, out = s.out;
  const int co1 = r * H1, co2 = r * H2;
  const bool mirror = (s.flags & 1) != 0;   // also store the intermediates to global memory
  const bool tc = (s.flags & 4) != 0;       // dense products on tensor-core tiles (mma.sync 3xTF32, tc_tiles.cuh)
  // flag bit 6, "head v2" (same arithmetic, same bits; fewer barriers and less gradient-row traffic):
  //   * the read-out row goes to global memory AFTER the cluster barrier of the exchange (its releasing arrive would
  //     otherwise wait for those stores);
  //   * regression with one output: every warp evaluates fc2, so prediction, loss term and dLoss/dpred live in every
  //     thread's registers - no barrier between fc2, the loss and the head backward, no single-thread section;
  //   * in-kernel reduction only: the fc1.weight gradient rows (Hd x C2 = 77 % of a gradient row) are NOT stored - the
  //     reducing CTA forms dh_g[j] * R_g[c] from the fc1.bias gradient row and the read-out row of graph g
  //     (rounded product, then the same four ordered quarter sums: bit-identical to summing stored rows)
  const bool head2 = (s.flags & 64) != 0;
  S2_PHASE(0);
  float* xs = sm + P.xs;   float* ax = sm + P.ax;   float* z1 = sm + P.z1;   float* dz1 = sm + P.dz1;
  float* w1t = sm + P.w1t; float* w2t = sm + P.w2t; float* w2 = sm + P.w2;
  float* p1 = sm + P.p1;   float* ap = sm + P.ap;   float* z2 = sm + P.z2;   float* p2 = sm + P.p2;
  float* dz2 = sm + P.dz2; float* dap = sm + P.dap; float* dp1 = sm + P.dp1;
  float* fc1w = sm + P.fc1w; float* fc2w = sm + P.fc2w; float* fc1b = sm + P.fc1b; float* fc2b = sm + P.fc2b;
  float* rrow = sm + P.rrow; float* hrow = sm + P.hrow; float* dhrow = sm + P.dhrow; float* drrow = sm + P.drrow;
  float* prow = sm + P.prow; float* red = sm + P.red;
  int* ism = reinterpret_cast<int*>(sm);
  int* arg0 = ism + P.arg0;  int* arg1 = ism + P.arg1;
  int* blb = ism + P.blob;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ism + P.bars);
  const int LDX = P.ldx, LDZ1 = P.ldz1, LDP = P.ldp, LDZ2 = P.ldz2;

  // ---- this branch's conv weights, requested before anything else (they do not depend on the graph): 4-byte
  // asynchronous copies whose DESTINATION address does the transposition, so no register, no dependent store and
  // ONE L2 round trip that overlaps the extent loads and the bulk-copy issue below (first commit group)
  {
#pragma unroll 1
    for (int i = t; i < H1 * F; i += T) {        // W1 [C1][F] rows co1.. (contiguous) -> w1t [F][H1]
      const int c = i / F, f = i - c * F;
      s2_cp4(w1t + f * H1 + c, a.W1 + (int64_t)co1 * F + i);
    }
#pragma unroll 1
    for (int i = t; i < H2 * H1; i += T) {       // W2 [2][H2][H1] group r -> w2 [H2][H1], w2t [H1][H2]
      const int o = i / H1, j = i - o * H1;
      const float* src = a.W2 + (int64_t)r * H2 * H1 + i;
      s2_cp4(w2 + i, src);
      s2_cp4(w2t + j * H2 + o, src);
    }
    s2_commit();
  }
  // ---- graph extents: ONE level of tiny loads (host-built pointers), then three bulk copies
  int n0, n, eg0, m;
  if (s.gdesc) {   // the record the structure pass left in L2 a moment ago: [K0, E1, K1, n0 | e0, m, 0, n]
    const int4 lo = __ldg(reinterpret_cast<const int4*>(s.gdesc + 8 * (int64_t)g));
    const int4 hi = __ldg(reinterpret_cast<const int4*>(s.gdesc + 8 * (int64_t)g + 4));
    n0 = lo.w; eg0 = hi.x; m = hi.y; n = hi.w;
  } else {
    n0 = __ldg(a.node_ptr + g); n = __ldg(a.node_ptr + g + 1) - n0;
    eg0 = __ldg(s.edge_ptr + g); m = __ldg(s.edge_ptr + g + 1) - eg0;
  }
  S2_PHASE(20);   // staging sub-phases 20..25 (diagnostic): extents known
  const bool train = !(s.forward_only || s.task == 0);
  float* part = s.partial + (int64_t)g * s.partial_ld;
  // loop-invariant global scalars of the head and this branch's weights, fetched while the copies fly
  const uint32_t drop_ctr = (!s.keep && s.drop_p > 0.f) ? (uint32_t)__ldg(s.step_dev) : 0u;
  float y_first = 0.f;
  int y_cls = 0;
  if (train && (t == 0 || head2)) {
    if (s.task == 3) y_cls = (int)__ldg(s.y_class + g);
    else y_first = __ldg(s.y + (int64_t)g * out);
  }
  const bool compact_fc1 = head2 && train && P.fused_reduce != 0;
  // The per-graph work sits in a do { } while (0): an invalid graph (host bounds violated, incomplete blob)
  // zeroes its gradient row, flags the status word and BREAKS to the grid barrier below instead of leaving the
  // kernel - every CTA must reach the barrier and take its ticket, or the counters would not be re-armed
  // for the next launch.  Both CTAs of a cluster see the same extents and take the same path.
  do {
  if (n < 0 || m < 0 || n > a.max_n || m > s.max_e) {   // host bounds violated: flag, contribute nothing (validate())
    s2_wait<0>();                                         // the weight copies issued above
    if (t == 0) atomicOr(a.status, 64);
    if (train && r == 0) {
#pragma unroll 1
      for (int i = t; i < s.n_params + 1; i += T) part[i] = 0.f;
#pragma unroll 1
      for (int i = t; i < C2; i += T) a.R[(int64_t)g * C2 + i] = 0.f;   // (head v2 multiplies it with the zero dh row)
    }
    break;
  }
  if (t == 0) {
    s2_mbar_init(&bars[0], 1);
    s2_mbar_init(&bars[1], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the initialised barriers, seen by the async proxy
    // the feature tile - or, when the structure pass left AX = A x ready (s.zin1, row stride LDX), those rows
    const uint32_t bbytes = (uint32_t)DRGNN_BLOB_USED(n, m) * 4u, xbytes = (uint32_t)(n * (s.zin1 ? LDX : F)) * 4u;
    s2_mbar_expect_tx(&bars[0], bbytes + xbytes);
    s2_bulk_g2s(blb, s.blob + DRGNN_BLOB_OFFSET(g, n0, eg0), bbytes, &bars[0]);
    if (xbytes) {
      if (s.zin1) s2_bulk_g2s(ax, s.zin1 + (int64_t)n0 * LDX, xbytes, &bars[0]);
      else s2_bulk_g2s(xs, a.x + (int64_t)n0 * F, xbytes, &bars[0]);
    }
    const uint32_t wbytes = (uint32_t)(Hd * C2) * 4u;
    s2_mbar_expect_tx(&bars[1], wbytes);
    s2_bulk_g2s(fc1w, s.fc1_w, wbytes, &bars[1]);
  }
  S2_PHASE(21);   // bulk copies issued
  // small head vectors (any alignment): cp.async
  s2_stage32(fc2w, s.fc2_w, out * Hd, t, T);
  if (s.fc1_b) s2_stage32(fc1b, s.fc1_b, Hd, t, T);
  if (s.fc2_b) s2_stage32(fc2b, s.fc2_b, out, t, T);
  s2_commit();
  S2_PHASE(22);   // head-vector copies issued
  if (!s.fc1_b) {
#pragma unroll 1
    for (int i = t; i < Hd; i += T) fc1b[i] = 0.f;
  }
  if (!s.fc2_b) {
#pragma unroll 1
    for (int i = t; i < out; i += T) fc2b[i] = 0.f;
  }
  s2_wait<1>();                  // this thread's conv-weight copies (the head vectors may still be in flight)
  S2_PHASE(23);
  __syncthreads();               // everybody's conv weights; barrier initialisation visible to every thread
  S2_PHASE(24);
  s2_mbar_wait(&bars[0], 0);     // structure blob + feature tile have landed
  S2_PHASE(25);
  const int K = blb[2], E1 = blb[3], Q = blb[4];
  if (blb[5] != 1 || blb[0] != n || blb[1] != m || K > a.max_k || Q > a.max_q || K < 0 || Q < 0 || E1 < 0 || E1 > m) {
    s2_mbar_wait(&bars[1], 0);   // no bulk copy may be in flight into a CTA that exits
    s2_wait<0>();
    if (t == 0) atomicOr(a.status, 64);
    if (train && r == 0) {
#pragma unroll 1
      for (int i = t; i < s.n_params + 1; i += T) part[i] = 0.f;
#pragma unroll 1
      for (int i = t; i < C2; i += T) a.R[(int64_t)g * C2 + i] = 0.f;   // (head v2 multiplies it with the zero dh row)
    }
    break;
  }
  const BlobLayout BL = blob_layout(n, m);
  const int* rp0 = blb + BL.rp0;     const int* col0 = blb + BL.col0;   const int* rp1 = blb + BL.rp1;  const int* col1 = blb + BL.col1;
  const int* cmp0 = blb + BL.cmp0;   const int* cmem0 = blb + BL.cmem0; const int* cl0 = blb + BL.cl0;
  const int* cmp1 = blb + BL.cmp1;   const int* cmem1 = blb + BL.cmem1; const int* cl1 = blb + BL.cl1;
  const int* cscp1 = blb + BL.cscp1; const int* cscr1 = blb + BL.cscr1;
  // The peer CTA must be running before its shared memory is written (read-out exchange below):
  // arrive now, wait just before the exchange, so the barrier costs nothing.
  // (relaxed: the barrier only tells the peer that this CTA runs - a releasing arrive would first wait for the
  // asynchronous copies of the head vectors, which are not needed before the read-out)
  asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  S2_PHASE(1);

  const int F4 = F >> 2, H14 = H1 >> 2, H24 = H2 >> 2;
  // ---- AX = A x : the F/4 lanes of a row share its edge list and read whole 16-byte-aligned feature rows
  if (!s.zin1) {
    s2_gather(rp0, col0, 0, 0, xs, F, ax, LDX, n, F4, t, T);
    __syncthreads();
  }
  S2_PHASE(2);
  // ---- Z1 = relu(AX W1_r^T)
  if (tc) tc_gemm(ax, LDX, w1t, H1, n, H1, F, z1, LDZ1, nullptr, 1, nullptr, 0, t, T);
  else s2_gemm(ax, LDX, w1t, H1, n, H1, F, z1, LDZ1, 1, t, T);
  __syncthreads();
  S2_PHASE(3);
  if (s.flags & 16) {   // diagnostic: the same product once more (warm instruction cache, same data) - clocks 26 / 27
    S2_PHASE(26);
    if (tc) tc_gemm(ax, LDX, w1t, H1, n, H1, F, z1, LDZ1, nullptr, 1, nullptr, 0, t, T);
    else s2_gemm(ax, LDX, w1t, H1, n, H1, F, z1, LDZ1, 1, t, T);
    __syncthreads();
    S2_PHASE(27);
    if (!s.zin1) s2_gather(rp0, col0, 0, 0, xs, F, ax, LDX, n, F4, t, T);
    __syncthreads();
    S2_PHASE(28);
  }
  // ---- P1 = cluster max of Z1 (community_pooling.py:201)
  s2_cluster_max(cmp0, cmem0, 0, 0, z1, LDZ1, p1, LDP, arg0, H1, K, H14, t, T);
  __syncthreads();
  S2_PHASE(4);
  // ---- AP = A1 P1 on the coarsened graph
  s2_gather(rp1, col1, 0, 0, p1, LDP, ap, LDP, K, H14, t, T);
  __syncthreads();
  S2_PHASE(5);
  // ---- Z2 = relu(AP W2_r^T)
  if (tc) tc_gemm(ap, LDP, w2t, H2, K, H2, H1, z2, LDZ2, nullptr, 1, nullptr, 0, t, T);
  else s2_gemm(ap, LDP, w2t, H2, K, H2, H1, z2, LDZ2, 1, t, T);
  __syncthreads();
  S2_PHASE(6);
  // ---- P2 = level-1 cluster max (max_pool_x)
  s2_cluster_max(cmp1, cmem1, 0, 0, z2, LDZ2, p2, H2, arg1, H2, Q, H24, t, T);
  __syncthreads();
  S2_PHASE(7);
  if (mirror) {   // parity tests: the intermediates the single-CTA kernel leaves in global memory (global ids)
    const int k0 = __ldg(a.kptr0 + g), q0 = __ldg(a.kptr1 + g);
    if (r == 0) s2_mirror(ax, LDX, a.Zin1 + (int64_t)n0 * F, F, n, F4, 0, t, T);
    s2_mirror(z1, LDZ1, a.Z1 + (int64_t)n0 * C1 + co1, C1, n, H14, 0, t, T);
    s2_mirror(arg0, H1, a.arg0 + (int64_t)k0 * C1 + co1, C1, K, H14, n0 + 1, t, T);
    s2_mirror(ap, LDP, a.Zin2 + (int64_t)k0 * C1 + co1, C1, K, H14, 0, t, T);
    s2_mirror(z2, LDZ2, a.Z2 + (int64_t)k0 * C2 + co2, C2, K, H24, 0, t, T);
    s2_mirror(arg1, H2, a.arg1 + (int64_t)q0 * C2 + co2, C2, Q, H24, k0 + 1, t, T);
  }
  // ---- R[g] half = mean over the graph's level-1 clusters; both halves land in both CTAs (DSMEM)
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  {
    float* peer_rrow = cluster.map_shared_rank(rrow, (unsigned)(r ^ 1));
#pragma unroll 1
    for (int c = t; c < H2; c += T) {
      float acc = 0.f;
      for (int q = 0; q < Q; ++q) acc += p2[q * H2 + c];
      acc *= 1.f / (float)max(Q, 1);
      if (!head2) a.R[(int64_t)g * C2 + co2 + c] = acc;
      rrow[co2 + c] = acc;
      peer_rrow[co2 + c] = acc;
    }
  }
  s2_wait<0>();                // small head vectors (this thread's copies) ...
  s2_mbar_wait(&bars[1], 0);   // ... and fc1.weight have landed
  cluster.sync();              // everybody's copies, and the peer's half of the read-out row
  if (head2) {
#pragma unroll 1
    for (int c = t; c < H2; c += T) a.R[(int64_t)g * C2 + co2 + c] = rrow[co2 + c];
  }
  S2_PHASE(8);
  // ---- fc1 (both CTAs, identical results): four lanes per hidden unit, each a quarter of the read-out
  // channels as 16-byte loads, two shuffles to combine
  {
    const int sub = t & 3;
#pragma unroll 1
    for (int j = t >> 2; j < Hd; j += T >> 2) {
      const float* wrow = fc1w + j * C2;
      float av = 0.f;
#pragma unroll 2
      for (int c = sub * 4; c < C2; c += 16) {
        const float4 wv = *reinterpret_cast<const float4*>(wrow + c);
        const float4 rv = *reinterpret_cast<const float4*>(rrow + c);
        av = fmaf(rv.x, wv.x, av); av = fmaf(rv.y, wv.y, av); av = fmaf(rv.z, wv.z, av); av = fmaf(rv.w, wv.w, av);
      }
      av += __shfl_xor_sync(0xffffffffu, av, 1);
      av += __shfl_xor_sync(0xffffffffu, av, 2);
      if (sub == 0) {
        float v = av + fc1b[j];
        v = v < 0.f ? 0.f : v;
        if (s.keep) {
          v = s.keep[(int64_t)g * Hd + j] > 0.f ? v * s.keep_scale : 0.f;
        } else if (s.drop_p > 0.f) {
          v = hash_uniform(s.seed, drop_ctr, (uint32_t)(g * Hd + j)) >= s.drop_p ? v * s.keep_scale : 0.f;
        }
        hrow[j] = v;
      }
    }
  }
  __syncthreads();
  S2_PHASE(9);
  const int j0 = r ? (Hd >> 1) : 0, j1 = r ? Hd : (Hd >> 1), nj = j1 - j0;   // hidden units whose gradient rows CTA r writes
  if (head2 && train && out == 1 && s.task == 1) {
    // ---- fc2, loss term, dLoss/dpred in EVERY warp (same lanes, same order: the same bits in every thread)
    float pv = 0.f;
#pragma unroll 1
    for (int j = lane; j < Hd; j += 32) pv = fmaf(hrow[j], fc2w[j], pv);
    pv = warp_sum(pv);
    pv += fc2b[0];
    const float dd = pv - y_first;
    const float dpred = 2.f * dd * s.inv_norm;
    if (t == 0) {
      prow[0] = dpred;
      if (r == 0) {
        s.pred[(int64_t)g * out] = pv;
        part[s.n_params] = (dd * dd) * s.inv_norm;   // summed into the loss by the reduction
        part[s.off_fc2b] = dpred;
      }
    }
    S2_PHASE(10);
    // ---- head backward: dh (all units, both CTAs); gradient slots of the hidden units [j0, j1) by CTA r
#pragma unroll 1
    for (int j = t; j < Hd; j += T) {
      const float hv = hrow[j];
      float acc = fmaf(dpred, fc2w[j], 0.f);
      acc = hv > 0.f ? acc * s.keep_scale : 0.f;
      dhrow[j] = acc;
      if (j >= j0 && j < j1) {
        part[s.off_fc1b + j] = acc;
        part[s.off_fc2w + j] = dpred * hv;
      }
    }
    __syncthreads();
  } else {
  // ---- fc2: warp per output
#pragma unroll 1
  for (int o = warp; o < out; o += NW) {
    float acc = 0.f;
#pragma unroll 1
    for (int j = lane; j < Hd; j += 32) acc = fmaf(hrow[j], fc2w[o * Hd + j], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += fc2b[o];
      prow[o] = acc;
      if (r == 0) s.pred[(int64_t)g * out + o] = acc;
    }
  }
  __syncthreads();
  S2_PHASE(10);
  if (!train) return;
  // ---- loss term of this graph and dLoss/dpred (one thread per CTA, identical results)
  if (t == 0) {
    float lg = 0.f;
    if (s.task == 3) {
      float mx = prow[0];
#pragma unroll 1
      for (int c = 1; c < out; ++c) mx = fmaxf(mx, prow[c]);
      float se = 0.f;
#pragma unroll 1
      for (int c = 0; c < out; ++c) se += expf(prow[c] - mx);
      const float lse = mx + logf(se);
      const int tc = y_cls;
      const float w = s.class_w ? s.class_w[tc] : 1.f;
      lg = w * (lse - prow[tc]);
#pragma unroll 1
      for (int c = 0; c < out; ++c) prow[c] = w * (expf(prow[c] - lse) - (c == tc ? 1.f : 0.f)) * s.inv_norm;
    } else {
#pragma unroll 1
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
#pragma unroll 1
  for (int j = t; j < Hd; j += T) {
    float acc = 0.f;
#pragma unroll 1
    for (int o = 0; o < out; ++o) acc = fmaf(prow[o], fc2w[o * Hd + j], acc);
    acc = hrow[j] > 0.f ? acc * s.keep_scale : 0.f;
    dhrow[j] = acc;
    if (j >= j0 && j < j1) part[s.off_fc1b + j] = acc;
  }
#pragma unroll 1
  for (int i = t; i < out * nj; i += T) {
    const int o = i / nj, j = j0 + (i - o * nj);
    part[s.off_fc2w + o * Hd + j] = prow[o] * hrow[j];
  }
  if (r == 0)
  {
#pragma unroll 1
    for (int o = t; o < out; o += T) part[s.off_fc2b + o] = prow[o];
  }
  __syncthreads();
  }
  if (!compact_fc1) {   // fc1.weight gradient rows of this CTA's hidden units: dh[j] * R[g][:], 16-byte stores
    const int C24 = C2 >> 2;
    const S2Div dv(C24);
#pragma unroll 1
    for (int i = t; i < nj * C24; i += T) {
      int jj, c4;
      dv.split(i, jj, c4);
      const float dh = dhrow[j0 + jj];
      float4 rv = *reinterpret_cast<const float4*>(rrow + c4 * 4);
      rv.x *= dh; rv.y *= dh; rv.z *= dh; rv.w *= dh;
      *reinterpret_cast<float4*>(part + s.off_fc1w + (j0 + jj) * C2 + c4 * 4) = rv;
    }
  }
  // dR[c] of this branch's channels = sum_j dh[j] W1[j][co2 + c]: warps split the hidden units, fixed-order sum
#pragma unroll 1
  for (int c = lane; c < H2; c += 32) {
    float acc = 0.f;
#pragma unroll 2
    for (int j = warp; j < Hd; j += NW) acc = fmaf(dhrow[j], fc1w[j * C2 + co2 + c], acc);
    red[warp * H2 + c] = acc;
  }
  __syncthreads();
#pragma unroll 1
  for (int c = t; c < H2; c += T) {
    float acc = 0.f;
#pragma unroll 4
    for (int w = 0; w < NW; ++w) acc += red[w * H2 + c];
    drrow[c] = acc;
  }
  __syncthreads();
  S2_PHASE(11);
  // ---- dZ2: read-out mean backward, routed to the arg-max member, gated by ReLU
  s2_route(cl1, 0, arg1, H2, z2, LDZ2, drrow, 0, 1.f / (float)max(Q, 1), dz2, LDZ2, K, H24, 0, t, T);
  __syncthreads();
  S2_PHASE(12);
  // ---- dW2_r = dZ2^T AP (split over the K0 rows, first half of the CTA)  ||  dAP = dZ2 W2_r (second half)
  const int KS2 = s2_split(P.xs_words, H2 * H1), KS1 = s2_split(P.xs_words, H1 * F);
  float* scratch = xs;   // the feature tile is dead since AX
  if (tc) {
    if (t < (T >> 1)) tc_splitk_partial(dz2, LDZ2, ap, LDP, H2, H1, K, KS2, scratch, t, T >> 1);
    else tc_gemm(dz2, LDZ2, w2, H1, K, H1, H2, dap, LDP, nullptr, 0, nullptr, 0, t - (T >> 1), T >> 1);
  } else {
    if (t < (T >> 1)) s2_splitk_partial(dz2, LDZ2, ap, LDP, H2, H1, K, KS2, scratch, t, T >> 1);
    else s2_gemm(dz2, LDZ2, w2, H1, K, H1, H2, dap, LDP, 0, t - (T >> 1), T >> 1);
  }
  __syncthreads();
  S2_PHASE(13);
  s2_splitk_reduce(scratch, H2 * H1, KS2, part + s.off_w2 + r * H2 * H1, t, T);
  // ---- dP1 = A1^T dAP  (CSC of the coarsened graph)
  s2_gather(cscp1, cscr1, 0, 0, dap, LDP, dp1, LDP, K, H14, t, T);
  __syncthreads();
  S2_PHASE(14);
  // ---- dZ1: routed to the arg-max node of its cluster, gated by ReLU
  s2_route(cl0, 0, arg0, H1, z1, LDZ1, dp1, LDP, 1.f, dz1, LDZ1, n, H14, 0, t, T);
  __syncthreads();
  S2_PHASE(15);
  // ---- dW1_r [H1][F] = dZ1^T AX, split over the nodes
  if (tc) tc_splitk_partial(dz1, LDZ1, ax, LDX, H1, F, n, KS1, scratch, t, T);
  else s2_splitk_partial(dz1, LDZ1, ax, LDX, H1, F, n, KS1, scratch, t, T);
  __syncthreads();
  s2_splitk_reduce(scratch, H1 * F, KS1, part + s.off_w1 + r * H1 * F, t, T);
  S2_PHASE(16);
  } while (0);
  if (!P.fused_reduce || !train) return;
  // ---- gradient reduction (+ Adam) inside this launch: the grid is co-resident (2B <= SMs, one CTA
  // per SM), so a grid barrier is safe; afterwards CTA c owns a slice of the flat gradient buffer and
  // sums the per-graph rows in four contiguous quarters (ascending), combined in ascending order.
  __syncthreads();
  unsigned* sync_ctr = reinterpret_cast<unsigned*>(s.step_dev + 2);
  if (t == 0) {
    __threadfence();
    atomicAdd(sync_ctr, 1u);
    const unsigned G = gridDim.x;
    const unsigned long long t0 = s2_globaltimer();
    while (s2_ld_acquire(sync_ctr) < G) {
      if (s2_globaltimer() - t0 > 2000000000ull) {   // 2 s (a time-sliced or instrumented GPU is slow, not wrong): flag and go on
        atomicOr(a.status, 128);
        break;
      }
    }
  } else if (t == 32) {
    // optimiser step / exchange epoch of this launch (by another warp, while thread 0 waits at the barrier),
    // read BEFORE the CTA's ticket is taken (below): the last ticket holder changes them
    if (s.fuse_adam) {
      const float st = *reinterpret_cast<volatile float*>(s.step_dev) + 1.f;
      red[0] = st;
      red[1] = adam_bias_correction(s.beta1, st);
      red[2] = adam_bias_correction(s.beta2, st);
    }
    if (C.world > 1) *reinterpret_cast<unsigned*>(red + 4) = *reinterpret_cast<volatile uint32_t*>(C.ctr) + 1u;
    __threadfence();
  }
  __syncthreads();
  S2_PHASE(17);
  {
    const int n = s.n_params, B = a.B;
    const int per = (n + 1 + (int)gridDim.x - 1) / (int)gridDim.x;   // elements of this CTA, 128 per sweep
    const int el = t & 127, q = t >> 7;
    float* psum = xs;   // [4][128]
    float* adamc = red; // [3] + exchange epoch
    unsigned* epoch_s = reinterpret_cast<unsigned*>(red + 4);
    const int world = C.world, rank = C.rank;
    const bool peers = world > 1;
    if (t == T - 1) {
      // Bookkeeping of the launch by a thread that has no element to reduce (per < 128), so its global round
      // trips overlap the loads of the others: the last CTA to take a ticket re-arms the grid barrier,
      // publishes the exchange epoch and bumps the optimiser step.  A CTA takes its ticket AFTER it passed the
      // barrier and after thread 0 read the step / epoch (fence + CTA barrier above), so nothing of this
      // launch can still need the old values when they change.
      unsigned* ticket = reinterpret_cast<unsigned*>(s.step_dev + 1);
      if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
        __threadfence();
        *ticket = 0u;
        *sync_ctr = 0u;
        if (peers) *reinterpret_cast<volatile uint32_t*>(C.ctr) = *epoch_s;
        if (s.fuse_adam) s.step_dev[0] = adamc[0];
      }
    }
    // ---- local sums of this CTA's slice; one GPU: Adam right away, several: delivered to every rank
#pragma unroll 1
    for (int sweep = 0; sweep < per; sweep += 128) {
      const int e = (int)blockIdx.x * per + sweep + el;
      const bool mine = sweep + el < per && e <= n;
      float acc = 0.f;
      // optimiser state of this element, fetched alongside the gradient rows (one L2 round trip, not two)
      float adam_mi = 0.f, adam_vi = 0.f, adam_pi = 0.f;
      if (q == 0 && mine && !peers && s.fuse_adam && e < n) {
        adam_mi = __ldcg(s.adam_m + e); adam_vi = __ldcg(s.adam_v + e); adam_pi = __ldcg(s.adam_p + e);
      }
      const int fe = e - s.off_fc1w;
      if (mine && compact_fc1 && fe >= 0 && fe < Hd * C2) {
        // fc1.weight element (j, c): sum over the graphs of dh_g[j] * R_g[c], the product rounded like the stored row
        // element would have been, the sums in the same order as below
        const int gs = (B + 3) >> 2;
        const int g0 = q * gs, g1 = min(B, g0 + gs);
        const int j = fe / C2, c = fe - j * C2;
        const float* dhp = s.partial + s.off_fc1b + j;
        const float* rp = a.R + c;
        int gg = g0;
#pragma unroll 1
        for (; gg + 16 <= g1; gg += 16) {
          float v[16], w[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) {
            v[u] = __ldcg(dhp + (int64_t)(gg + u) * s.partial_ld);
            w[u] = __ldcg(rp + (int64_t)(gg + u) * C2);
          }
#pragma unroll
          for (int u = 0; u < 16; ++u) acc += __fmul_rn(w[u], v[u]);
        }
#pragma unroll 1
        for (; gg < g1; ++gg) acc += __fmul_rn(__ldcg(rp + (int64_t)gg * C2), __ldcg(dhp + (int64_t)gg * s.partial_ld));
      } else if (mine) {
        const int gs = (B + 3) >> 2;
        const int g0 = q * gs, g1 = min(B, g0 + gs);
        const float* src = s.partial + e;
        int gg = g0;
#pragma unroll 1
        for (; gg + 16 <= g1; gg += 16) {     // 16 independent loads in flight (the whole quarter at batch 64)
          float v[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) v[u] = __ldcg(src + (int64_t)(gg + u) * s.partial_ld);
#pragma unroll
          for (int u = 0; u < 16; ++u) acc += v[u];
        }
#pragma unroll 1
        for (; gg + 8 <= g1; gg += 8) {
          float v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = __ldcg(src + (int64_t)(gg + u) * s.partial_ld);
#pragma unroll
          for (int u = 0; u < 8; ++u) acc += v[u];
        }
#pragma unroll 1
        for (; gg < g1; ++gg) acc += __ldcg(src + (int64_t)gg * s.partial_ld);
      }
      psum[q * 128 + el] = acc;
      __syncthreads();
      if (q == 0 && mine) {
        acc = ((psum[el] + psum[128 + el]) + psum[256 + el]) + psum[384 + el];
        if (peers) {
          // ONE 8-byte store {value, epoch} per rank over NVLink peer memory into slot `rank` of every
          // rank's low-latency buffer (own included): validity travels with the value, no flag, no fence
          const unsigned ep = *epoch_s;
          const int64_t slot = ((int64_t)(ep & 1u) * world + rank) * C.stride + e;
#pragma unroll 1
          for (int p = 0; p < world; ++p) s2_st_ll(C.xll[(rank + p) % world] + slot, __float_as_uint(acc), ep);
        } else if (e < n) {
          s.grads[e] = acc;
          if (s.fuse_adam) {
            float mi = adam_mi, vi = adam_vi;
            mi = mi + (acc - mi) * (1.f - s.beta1);
            vi = vi * s.beta2 + (1.f - s.beta2) * acc * acc;
            s.adam_m[e] = mi;
            s.adam_v[e] = vi;
            const float denom = sqrtf(vi) / sqrtf(adamc[2]) + s.eps;
            s.adam_p[e] = adam_pi - (s.lr / adamc[1]) * (mi / denom);
          }
        } else if (s.loss) {
          s.loss[0] = acc;
        }
      }
      __syncthreads();
    }
    if (peers) {
      // ---- every thread waits for ITS element from every rank (element e only depends on element e of
      // the peers: no flag, no cross-GPU barrier), sums the slots in rank order and applies Adam
      const unsigned epoch = *epoch_s;
      const int par = (int)(epoch & 1u);
      const unsigned long long t0 = s2_globaltimer(), limit = C.timeout_ns ? C.timeout_ns : 20000000000ull;
#pragma unroll 1
      for (int sweep = 0; sweep < per; sweep += T) {
        const int e = (int)blockIdx.x * per + sweep + t;
        if (sweep + t < per && e <= n) {
          float tot = 0.f;
          const uint64_t* mine = C.xll[rank] + (int64_t)par * world * C.stride + e;
#pragma unroll 1
          for (int rr = 0; rr < world; ++rr) {
            unsigned bits, ep;
            s2_ld_ll(mine + (int64_t)rr * C.stride, bits, ep);
            while (ep != epoch) {
              if (s2_globaltimer() - t0 > limit) {
                atomicOr(C.ctr + 2, 1u);
                break;
              }
              __nanosleep(20);
              s2_ld_ll(mine + (int64_t)rr * C.stride, bits, ep);
            }
            tot += __uint_as_float(bits);
          }
          if (e < n) {
            s.grads[e] = tot;
            if (s.fuse_adam) {
              float mi = s.adam_m[e], vi = s.adam_v[e];
              mi = mi + (tot - mi) * (1.f - s.beta1);
              vi = vi * s.beta2 + (1.f - s.beta2) * tot * tot;
              s.adam_m[e] = mi;
              s.adam_v[e] = vi;
              const float denom = sqrtf(vi) / sqrtf(adamc[2]) + s.eps;
              s.adam_p[e] = s.adam_p[e] - (s.lr / adamc[1]) * (mi / denom);
            }
          } else if (s.loss) {
            s.loss[0] = tot;
          }
        }
      }
    }
  }
  S2_PHASE(18);
}

static inline bool step2_shapes_ok(const drgnn_ginet_step_args& s) {
  const drgnn_ginet_fused_args& a = s.g;
  return a.nb == 2 && s.blob != nullptr && s.edge_ptr != nullptr && a.F % 4 == 0 && a.h1 % 4 == 0 && a.h2 % 4 == 0 && s.max_e > 0 && a.max_n > 0 && a.max_k > 0 &&
         a.max_q > 0 && s.Hd > 0 && s.out > 0 && ((2 * a.h2 * s.Hd) % 4 == 0) &&
         // 16-byte stores of the fc1.weight gradient rows (training only)
         (s.forward_only || s.task == 0 ||
          (s.partial_ld % 4 == 0 && s.off_fc1w % 4 == 0 && ((uintptr_t)s.partial % 16) == 0));
}
