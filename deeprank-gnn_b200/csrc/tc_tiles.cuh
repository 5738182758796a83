// Dense products of the fused per-graph step kernels on tensor-core tiles.  Included (inside namespace drgnn) by
// fused.cu (CTA-pair GINet kernel) and step3.cu (general cluster kernel).
#pragma once

// ---------------------------------------------------------------------------------------------------------
// The same two dense products on the tensor cores: mma.sync.m16n8k8 TF32 with the 3-product error compensation
// (a = a_hi + a_lo, b = b_hi + b_lo, d += a_lo b_hi + a_hi b_lo + a_hi b_hi; fp32 accumulate, error ~1e-6) - the
// "dense per-node feature x weight contraction" of north_star inside the fused step.  A warp owns 16 x 8 output
// tiles; operands are read from shared memory as fragments (gid = lane / 4, tig = lane % 4):
//   A 16x8: a0 (gid, tig) a1 (gid+8, tig) a2 (gid, tig+4) a3 (gid+8, tig+4);  B 8x8: b0 (k tig, n gid) b1 (k tig+4, n gid)
//   C 16x8: c0 (gid, 2 tig) c1 (gid, 2 tig+1) c2 (gid+8, 2 tig) c3 (gid+8, 2 tig+1)
// (a tcgen05 tile is not worth its round trip here: M <= 256 rows, N <= 64, K <= 64 per CTA and phase.)
__device__ __forceinline__ uint32_t tc_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void tc_split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = tc_tf32(x);
  lo = tc_tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void tc_mma_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// C[m][n] = act( sum_k A[m*lda + k] * Bm[k*ldb + n] + bias[n] ) (* rscale[m] for columns >= scol) on the tensor cores
// (the contract of s2_gemm / s3_gemm; K % 8 == 0, N % 8 == 0).  Rows past M of the last 16-row tile read
// whatever follows the operand in shared memory (an output row depends on its own input row only) and are not stored.
static __device__ __noinline__ void tc_gemm(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, int M, int N,
                                        int K, float* __restrict__ C, int ldc, const float* __restrict__ bias, int relu,
                                        const float* __restrict__ rscale, int scol, int tid, int nth) {
  const int warp = tid >> 5, nwarps = nth >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int mt = (M + 15) >> 4, nt = N >> 3;
#pragma unroll 1
  for (int item = warp; item < mt * nt; item += nwarps) {
    const int mi = item / nt, ni = item - mi * nt;
    const float* a0p = A + (mi * 16 + gid) * lda + tig;
    const float* a1p = a0p + 8 * lda;
    const float* bp = Bm + tig * ldb + ni * 8 + gid;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int k = 0; k < K; k += 8) {
      uint32_t ahi[4], alo[4], bhi[2], blo[2];
      tc_split_tf32(a0p[k], ahi[0], alo[0]);
      tc_split_tf32(a1p[k], ahi[1], alo[1]);
      tc_split_tf32(a0p[k + 4], ahi[2], alo[2]);
      tc_split_tf32(a1p[k + 4], ahi[3], alo[3]);
      tc_split_tf32(bp[k * ldb], bhi[0], blo[0]);
      tc_split_tf32(bp[(k + 4) * ldb], bhi[1], blo[1]);
      tc_mma_tf32(acc, alo, bhi);
      tc_mma_tf32(acc, ahi, blo);
      tc_mma_tf32(acc, ahi, bhi);
    }
    const int n0 = ni * 8 + 2 * tig;
    float b0 = 0.f, b1 = 0.f;
    if (bias) { b0 = bias[n0]; b1 = bias[n0 + 1]; }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int m = mi * 16 + gid + 8 * h;
      if (m < M) {
        float x = acc[2 * h] + b0, y = acc[2 * h + 1] + b1;
        if (relu) { x = x < 0.f ? 0.f : x; y = y < 0.f ? 0.f : y; }       // keeps NaN like torch.relu
        if (rscale && n0 >= scol) { const float r_ = rscale[m]; x *= r_; y *= r_; }
        *reinterpret_cast<float2*>(C + m * ldc + n0) = make_float2(x, y);
      }
    }
  }
}

// Split-K partial products (the contract of s2_splitk_partial / s3_splitk_partial) on the tensor cores: scratch[s][M][N] = sum over the k-range of split s of At[k][m] Bm[k][n]
// (N % 8 == 0; rows m >= M of a 16-row tile and k beyond the range contribute zeros / are not stored).
static __device__ __noinline__ void tc_splitk_partial(const float* __restrict__ At, int lda, const float* __restrict__ Bm, int ldb, int M,
                                                  int N, int K, int KS, float* __restrict__ scratch, int tid, int nth) {
  const int warp = tid >> 5, nwarps = nth >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int mt = (M + 15) >> 4, nt = N >> 3, tiles = mt * nt;
  const int chunk = (((K + KS - 1) / KS) + 7) & ~7;          // k-range of a split: a multiple of the MMA depth
#pragma unroll 1
  for (int item = warp; item < tiles * KS; item += nwarps) {
    const int sp = item / tiles, tile = item - sp * tiles;
    const int mi = tile / nt, ni = tile - mi * nt;
    const int kb = sp * chunk, ke = min(K, kb + chunk);
    const int m0 = mi * 16 + gid, m1 = m0 + 8;
    const bool v0 = m0 < M, v1 = m1 < M;
    const float* bp = Bm + ni * 8 + gid;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 2
    for (int k = kb; k < ke; k += 8) {
      const int ka = k + tig, kc = k + tig + 4;
      const bool ua = ka < ke, uc = kc < ke;
      uint32_t ahi[4], alo[4], bhi[2], blo[2];
      tc_split_tf32((ua && v0) ? At[ka * lda + m0] : 0.f, ahi[0], alo[0]);
      tc_split_tf32((ua && v1) ? At[ka * lda + m1] : 0.f, ahi[1], alo[1]);
      tc_split_tf32((uc && v0) ? At[kc * lda + m0] : 0.f, ahi[2], alo[2]);
      tc_split_tf32((uc && v1) ? At[kc * lda + m1] : 0.f, ahi[3], alo[3]);
      tc_split_tf32(ua ? bp[ka * ldb] : 0.f, bhi[0], blo[0]);
      tc_split_tf32(uc ? bp[kc * ldb] : 0.f, bhi[1], blo[1]);
      tc_mma_tf32(acc, alo, bhi);
      tc_mma_tf32(acc, ahi, blo);
      tc_mma_tf32(acc, ahi, bhi);
    }
    float* sp_ = scratch + (size_t)sp * M * N + ni * 8 + 2 * tig;
    if (v0) *reinterpret_cast<float2*>(sp_ + m0 * N) = make_float2(acc[0], acc[1]);
    if (v1) *reinterpret_cast<float2*>(sp_ + m1 * N) = make_float2(acc[2], acc[3]);
  }
}

