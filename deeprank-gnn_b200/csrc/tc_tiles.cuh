// Dense products of the fused per-graph step kernels on tensor-core tiles.  Included (inside namespace drgnn) by
// fused.cu (CTA-pair GINet kernel) and step3.cu (general cluster kernel).
#pragma once

// ---------------------------------------------------------------------------------------------------------
// The two dense products of the step kernels (row-major transform, split-K weight gradient) on the tensor cores:
// mma.sync.m16n8k8 TF32 with the 3-product error compensation (a = a_hi + a_lo, b = b_hi + b_lo,
// d = sum a_hi b_hi + (sum a_lo b_hi + sum a_hi b_lo), three independent fp32 accumulator chains; error ~1e-6) - the "dense per-node feature x weight
// contraction" of north_star inside the fused step.  Fragments (gid = lane / 4, tig = lane % 4):
//   A 16x8: a0 (gid, tig) a1 (gid+8, tig) a2 (gid, tig+4) a3 (gid+8, tig+4);  B 8x8: b0 (k tig, n gid) b1 (k tig+4, n gid)
//   C 16x8: c0 (gid, 2 tig) c1 (gid, 2 tig+1) c2 (gid+8, 2 tig) c3 (gid+8, 2 tig+1)
// The hi / lo split is ONE mask + ONE subtract per element: the tensor core reads only the upper 19 bits of a
// TF32 operand, so hi is the fp32 word itself and lo = x - (x & 0xffffe000) (exact) - cvt.rna.tf32 runs on the
// quarter-rate conversion pipe and made the first version of these tiles slower than the FFMA tiles.  A warp owns
// a 16 x (8 NTI) block of the output, so one A fragment feeds NTI independent accumulator chains; row-major A
// fragments come from ONE ldmatrix.x4 (a TF32 element = a pair of b16).
// (a tcgen05 tile is not worth its round trip here: M <= 256 rows, N <= 64, K <= 64 per CTA and phase.)
__device__ __forceinline__ void tc_split(float x, uint32_t& hi, uint32_t& lo) {
  const uint32_t xb = __float_as_uint(x);
  hi = xb;
  lo = __float_as_uint(x - __uint_as_float(xb & 0xffffe000u));
}
__device__ __forceinline__ void tc_mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void tc_ldsm4(uint32_t (&r)[4], const float* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"((uint32_t)__cvta_generic_to_shared(p))
               : "memory");
}

template <int NTI>
__device__ __forceinline__ void tc_gemm_items(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, int M,
                                              int N, int K, float* __restrict__ C, int ldc, const float* __restrict__ bias,
                                              int relu, const float* __restrict__ rscale, int scol, int tid, int nth) {
  const int warp = tid >> 5, nwarps = nth >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int mt = (M + 15) >> 4, ngr = (N >> 3) / NTI;
#pragma unroll 1
  for (int item = warp; item < mt * ngr; item += nwarps) {
    const int mi = item / ngr, ng = item - mi * ngr;
    // ldmatrix row of this lane: matrix lane / 8 = (rows +0 | +8) x (k +0 | +4); rows past M re-read row M - 1
    const int lrow = min(mi * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, M - 1);
    const float* ap = A + lrow * lda + (lane >> 4) * 4;
    const float* bp = Bm + tig * ldb + ng * (8 * NTI) + gid;
    float acc[NTI][4], acl[NTI][4], ach[NTI][4];    // hi.hi | lo.hi | hi.lo: three independent accumulator chains
#pragma unroll
    for (int j = 0; j < NTI; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = acl[j][i] = ach[j][i] = 0.f;
#pragma unroll 2
    for (int k = 0; k < K; k += 8) {
      uint32_t ar[4], ahi[4], alo[4];
      tc_ldsm4(ar, ap + k);
#pragma unroll
      for (int i = 0; i < 4; ++i) tc_split(__uint_as_float(ar[i]), ahi[i], alo[i]);
#pragma unroll
      for (int j = 0; j < NTI; ++j) {
        uint32_t b0h, b0l, b1h, b1l;
        tc_split(bp[k * ldb + 8 * j], b0h, b0l);
        tc_split(bp[(k + 4) * ldb + 8 * j], b1h, b1l);
        tc_mma_tf32(acl[j], alo, b0h, b1h);
        tc_mma_tf32(ach[j], ahi, b0l, b1l);
        tc_mma_tf32(acc[j], ahi, b0h, b1h);
      }
    }
#pragma unroll
    for (int j = 0; j < NTI; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] += acl[j][i] + ach[j][i];      // the small terms first
#pragma unroll
    for (int j = 0; j < NTI; ++j) {
      const int n0 = ng * (8 * NTI) + 8 * j + 2 * tig;
      float b0 = 0.f, b1 = 0.f;
      if (bias) { b0 = bias[n0]; b1 = bias[n0 + 1]; }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = mi * 16 + gid + 8 * h;
        if (m < M) {
          float x = acc[j][2 * h] + b0, y = acc[j][2 * h + 1] + b1;
          if (relu) { x = x < 0.f ? 0.f : x; y = y < 0.f ? 0.f : y; }       // keeps NaN like torch.relu
          if (rscale && n0 >= scol) { const float r_ = rscale[m]; x *= r_; y *= r_; }
          *reinterpret_cast<float2*>(C + m * ldc + n0) = make_float2(x, y);
        }
      }
    }
  }
}

// C[m][n] = act( sum_k A[m*lda + k] * Bm[k*ldb + n] + bias[n] ) (* rscale[m] for columns >= scol) on the tensor cores
// (the contract of s2_gemm / s3_gemm; K % 8 == 0, N % 8 == 0, rows of A 16-byte aligned).
static __device__ __noinline__ void tc_gemm(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, int M,
                                            int N, int K, float* __restrict__ C, int ldc, const float* __restrict__ bias, int relu,
                                            const float* __restrict__ rscale, int scol, int tid, int nth) {
  if (M <= 0) return;
  // 16 columns per warp item when that still gives every warp work, else 8
  const int nwarps = nth >> 5, mt = (M + 15) >> 4, nt = N >> 3;
  if ((nt & 1) == 0 && mt * (nt >> 1) >= nwarps) tc_gemm_items<2>(A, lda, Bm, ldb, M, N, K, C, ldc, bias, relu, rscale, scol, tid, nth);
  else tc_gemm_items<1>(A, lda, Bm, ldb, M, N, K, C, ldc, bias, relu, rscale, scol, tid, nth);
}

template <int NTI>
__device__ __forceinline__ void tc_splitk_items(const float* __restrict__ At, int lda, const float* __restrict__ Bm, int ldb, int M,
                                                int N, int K, int KS, float* __restrict__ scratch, int tid, int nth) {
  const int warp = tid >> 5, nwarps = nth >> 5, lane = tid & 31, gid = lane >> 2, tig = lane & 3;
  const int mt = (M + 15) >> 4, ngr = (N >> 3) / NTI, tiles = mt * ngr;
  const int chunk = (((K + KS - 1) / KS) + 7) & ~7;          // k-range of a split: a multiple of the MMA depth
#pragma unroll 1
  for (int item = warp; item < tiles * KS; item += nwarps) {
    const int sp = item / tiles, tile = item - sp * tiles;
    const int mi = tile / ngr, ng = tile - mi * ngr;
    const int kb = sp * chunk, ke = min(K, kb + chunk);
    const int m0 = mi * 16 + gid, m1 = m0 + 8;
    const bool v0 = m0 < M, v1 = m1 < M;
    const float* ap0 = At + min(m0, M - 1);                   // columns past M re-read column M - 1 (not stored)
    const float* ap1 = At + min(m1, M - 1);
    const float* bp = Bm + ng * (8 * NTI) + gid;
    float acc[NTI][4], acl[NTI][4], ach[NTI][4];    // hi.hi | lo.hi | hi.lo: three independent accumulator chains
#pragma unroll
    for (int j = 0; j < NTI; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = acl[j][i] = ach[j][i] = 0.f;
    const int kfull = kb + (max(ke - kb, 0) & ~7);
#pragma unroll 2
    for (int k = kb; k < kfull; k += 8) {
      const int ka = (k + tig) * lda, kc = (k + tig + 4) * lda;
      uint32_t ahi[4], alo[4];
      tc_split(ap0[ka], ahi[0], alo[0]);
      tc_split(ap1[ka], ahi[1], alo[1]);
      tc_split(ap0[kc], ahi[2], alo[2]);
      tc_split(ap1[kc], ahi[3], alo[3]);
      const float* b0p = bp + (k + tig) * ldb;
      const float* b1p = bp + (k + tig + 4) * ldb;
#pragma unroll
      for (int j = 0; j < NTI; ++j) {
        uint32_t b0h, b0l, b1h, b1l;
        tc_split(b0p[8 * j], b0h, b0l);
        tc_split(b1p[8 * j], b1h, b1l);
        tc_mma_tf32(acl[j], alo, b0h, b1h);
        tc_mma_tf32(ach[j], ahi, b0l, b1l);
        tc_mma_tf32(acc[j], ahi, b0h, b1h);
      }
    }
    if (kfull < ke) {                                         // tail of the k-range: rows past ke contribute zeros
      const int k = kfull;
      const bool ua = k + tig < ke, uc = k + tig + 4 < ke;
      const int ka = (k + tig) * lda, kc = (k + tig + 4) * lda;
      uint32_t ahi[4], alo[4];
      tc_split(ua ? ap0[ka] : 0.f, ahi[0], alo[0]);
      tc_split(ua ? ap1[ka] : 0.f, ahi[1], alo[1]);
      tc_split(uc ? ap0[kc] : 0.f, ahi[2], alo[2]);
      tc_split(uc ? ap1[kc] : 0.f, ahi[3], alo[3]);
#pragma unroll
      for (int j = 0; j < NTI; ++j) {
        uint32_t b0h, b0l, b1h, b1l;
        tc_split(ua ? bp[(k + tig) * ldb + 8 * j] : 0.f, b0h, b0l);
        tc_split(uc ? bp[(k + tig + 4) * ldb + 8 * j] : 0.f, b1h, b1l);
        tc_mma_tf32(acl[j], alo, b0h, b1h);
        tc_mma_tf32(ach[j], ahi, b0l, b1l);
        tc_mma_tf32(acc[j], ahi, b0h, b1h);
      }
    }
#pragma unroll
    for (int j = 0; j < NTI; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] += acl[j][i] + ach[j][i];      // the small terms first
#pragma unroll
    for (int j = 0; j < NTI; ++j) {
      float* sp_ = scratch + (size_t)sp * M * N + ng * (8 * NTI) + 8 * j + 2 * tig;
      if (v0) *reinterpret_cast<float2*>(sp_ + m0 * N) = make_float2(acc[j][0], acc[j][1]);
      if (v1) *reinterpret_cast<float2*>(sp_ + m1 * N) = make_float2(acc[j][2], acc[j][3]);
    }
  }
}

// Split-K partial products (the contract of s2_splitk_partial / s3_splitk_partial) on the tensor cores:
// scratch[s][M][N] = sum over the k-range of split s of At[k][m] Bm[k][n]  (N % 8 == 0).
static __device__ __noinline__ void tc_splitk_partial(const float* __restrict__ At, int lda, const float* __restrict__ Bm, int ldb,
                                                      int M, int N, int K, int KS, float* __restrict__ scratch, int tid, int nth) {
  if (M <= 0) return;
  const int nwarps = nth >> 5, mt = (M + 15) >> 4, nt = N >> 3;
  if ((nt & 3) == 0 && mt * (nt >> 2) * KS >= nwarps) tc_splitk_items<4>(At, lda, Bm, ldb, M, N, K, KS, scratch, tid, nth);
  else if ((nt & 1) == 0 && mt * (nt >> 1) * KS >= nwarps) tc_splitk_items<2>(At, lda, Bm, ldb, M, N, K, KS, scratch, tid, nth);
  else tc_splitk_items<1>(At, lda, Bm, ldb, M, N, K, KS, scratch, tid, nth);
}
