// Whole training / scoring step of ONE graph on a thread-block CLUSTER, for the three reference networks
// (GINet ginet.py:99-141, sGAT sGAT.py:114-138, FoutNet foutnet.py:103-125) and for graphs of any size the
// cluster's shared memory holds.  Included by step3.cu (inside namespace drgnn).
//
// ginet_graph_step2_kernel (fused_step2.cuh) is the tuned special case "GINet, one graph fits a CTA pair".
// This kernel generalises it along two axes:
//
//   * KIND: the aggregation that feeds the dense transform of a conv layer
//       GINet   zin_i = sum_{e: row=i} x_col                                  z = relu(zin W^T)        (alpha == 1)
//       sGAT    zin_i = [ s_i x_i | (1/max(deg_i,1)) sum_e a_e x_col ]        z = relu(zin W + b)
//                       s_i = (1/max(deg_i,1)) sum_e a_e                      (sGAT.py:70-92, factorised)
//       FoutNet zin_i = [ x_i | (1/deg_i) sum_e x_col ]  (deg_i = 0 -> NaN)   z = relu(zin [Wc;Wn] + b)  (foutnet.py:62-80)
//     GINet runs its two branches on two groups of CTAs of the cluster (as step2 does), the others on one.
//   * TILES: the node dimension of a graph is split over NT CTAs of the cluster (NT = 1, 2, 4, 8).  CTA `ti`
//     owns rows [ti*ceil(n/NT), ...) of the level-0 graph, of the coarsened graph and of the level-1 clusters;
//     every intermediate of its rows lives in ITS shared memory and the other CTAs of the cluster read the rows
//     they need (neighbour gathers, cluster members, arg-max routing) through DISTRIBUTED SHARED MEMORY
//     (cluster.map_shared_rank); a cluster barrier separates producer and consumer phases.  cfg4 (500 nodes,
//     hidden 32/64: 2 branches x 4 tiles) and cfg5 (up to 1000 nodes) then stay in ONE launch instead of ~25.
//     With NT = 1 the graph's structure blob and feature tile are staged by bulk copies (TMA) like step2; with
//     NT > 1 the index lists and level-0 features are read straight from global memory / L2 (each entry is used
//     once per CTA; the lists of a tile are not bounded by n/NT, so no shared-memory capacity can be promised).
//
// Per graph the phases are those of step2 (same fmaf order per output element when NT = 1), the weight
// gradients are split-K products over the CTA's own rows, summed over the tiles through DSMEM in tile order
// (deterministic), and the per-graph gradient rows are reduced behind a grid barrier inside the launch when the
// grid is co-resident (+ Adam, + the NVLink peer exchange), else by net_step_reduce_kernel.
#pragma once

namespace cgx = cooperative_groups;

static constexpr int S3_THREADS = 512;
static constexpr int S3_MAX_TILES = 8;
static constexpr int S3_MAX_KS = 16;

__device__ unsigned long long g_phase3[32];
// diagnostic (flag bit 3): %globaltimer of every CTA at kernel entry and at the end of its per-graph work
// (drgnn_debug_cta_times: start skew and load imbalance across the grid - what the grid barrier waits for)
__device__ unsigned long long g_cta_times[2048][2];
// (flag bit 3 of the launch enables the clocks: their global stores delay the release fences of block 0, and the
// whole grid waits for block 0 at the grid barrier)
#define DRGNN_PHASE3(i)                                                                       \
  do {                                                                                        \
    if (timers) g_phase3[i] = (unsigned long long)clock64();                                  \
  } while (0)

__device__ __forceinline__ uint32_t s3_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void s3_cp4(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s3_smem_u32(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void s3_mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s3_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void s3_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s3_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void s3_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s3_smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(s3_smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void s3_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "S3_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra S3_DONE_%=;\n"
      "bra S3_WAIT_%=;\n"
      "S3_DONE_%=:\n"
      "}\n" ::"r"(s3_smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ unsigned s3_ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void s3_st_ll(uint64_t* p, unsigned bits, unsigned epoch) {
  asm volatile("st.relaxed.sys.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(bits), "r"(epoch) : "memory");
}
__device__ __forceinline__ void s3_ld_ll(const uint64_t* p, unsigned& bits, unsigned& epoch) {
  asm volatile("ld.relaxed.sys.global.v2.u32 {%0, %1}, [%2];" : "=r"(bits), "=r"(epoch) : "l"(p) : "memory");
}
__device__ __forceinline__ unsigned long long s3_globaltimer() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}
__device__ __forceinline__ float s3_hash_uniform(uint32_t seed, uint32_t ctr, uint32_t idx) {   // == hash_uniform (fused.cu)
  uint32_t x = idx * 0x9E3779B1u ^ (ctr * 0x85EBCA77u) ^ (seed * 0xC2B2AE3Du);
  x ^= x >> 16; x *= 0x7FEB352Du;
  x ^= x >> 15; x *= 0x846CA68Bu;
  x ^= x >> 16;
  return (float)(x >> 8) * (1.0f / 16777216.0f);
}

__host__ __device__ inline int s3_up4(int x) { return (x + 3) & ~3; }
__host__ __device__ inline int s3_up8(int x) { return (x + 7) & ~7; }
__host__ __device__ inline int s3_cdiv(int a, int b) { return (a + b - 1) / b; }

// Shared-memory plan of one CTA (word offsets, every offset a multiple of 4 words = 16 bytes).
struct Step3Plan {
  int tiles, nbr, cs;             // node tiles, branches (2 for GINet), cluster size = tiles * nbr
  int nt, kt, qt;                 // row capacities of one tile at the three levels
  int Kin1, Kin2;                 // widths of the transform inputs: F | 2F, h1 | 2h1
  int ldzin1, ldz1, ldp, ldzin2, ldz2;
  int xs, zin1, z1, p1, zin2, z2, p2, wg, scr, w1, w2, w2t, b1, b2, fc2w, fc1b, fc2b;
  int zin3, z3, t3, w3, w3t, b3, ldzin3, layers3;     // optional third conv layer on the coarsened graph
  int rrow, hrow, dhrow, drrow, prow, red, rpart, s1, post1;
  int arg0, arg1, blob, wblob, bases, bars;
  int xs_words, scr_words, wg_words, blob_words;
  int total;
  int fused_reduce;
};

__host__ __device__ inline Step3Plan step3_plan(int kind, int tiles, int stage, int F, int h1, int h2, int max_n, int max_k,
                                                int max_q, int max_e, int Hd, int out, int layers3 = 0) {
  Step3Plan p;
  p.layers3 = (layers3 && kind != 0) ? 1 : 0;
  p.tiles = tiles;
  p.nbr = kind == 0 ? 2 : 1;
  p.cs = tiles * p.nbr;
  p.nt = s3_up8(s3_cdiv(max_n, tiles));
  p.kt = s3_up8(s3_cdiv(max_k, tiles));
  p.qt = s3_up8(s3_cdiv(max_q, tiles));
  p.Kin1 = kind == 0 ? F : 2 * F;
  p.Kin2 = kind == 0 ? h1 : 2 * h1;
  p.ldzin1 = p.Kin1 + 4; p.ldz1 = h1 + 4; p.ldp = h1 + 4; p.ldzin2 = p.Kin2 + 4; p.ldz2 = h2 + 4;
  const int C2 = p.nbr * h2;
  int o = 0;
  auto take = [&](int words) { const int at = o; o += s3_up4(words); return at; };
  // weight-gradient products: [M][N] with M = Cout (GINet, dZ^T zin) or Kin + 4 (the others, zin^T dZ with the
  // ones column of zin giving the bias gradient)
  const int mn1 = kind == 0 ? h1 * F : (p.Kin1 + 4) * h1;
  const int mn2 = kind == 0 ? h2 * h1 : (p.Kin2 + 4) * h2;
  const int mn3 = p.layers3 ? (2 * h2 + 4) * h2 : 0;
  p.wg_words = s3_up4(mn1 > mn2 ? (mn1 > mn3 ? mn1 : mn3) : (mn2 > mn3 ? mn2 : mn3));
  p.xs_words = stage ? s3_up4(max_n * F) : 0;
  int scr = 4 * p.wg_words;                       // at least four K splits
  if (scr < 4 * 128) scr = 4 * 128;               // partial sums of the in-kernel gradient reduction
  if (scr < p.xs_words) scr = p.xs_words;         // staged: the feature tile is dead after the first aggregation
  p.scr_words = scr;
  p.scr = take(scr);
  p.xs = p.scr;
  p.zin1 = take(p.nt * p.ldzin1);
  p.z1 = take(p.nt * p.ldz1);
  p.p1 = take(p.kt * p.ldp);
  p.zin2 = take(p.kt * p.ldzin2);
  p.z2 = take(p.kt * p.ldz2);
  p.p2 = take(p.qt * h2);
  p.ldzin3 = 2 * h2 + 4;
  p.zin3 = take(p.layers3 ? p.kt * p.ldzin3 : 0);
  p.z3 = take(p.layers3 ? p.kt * p.ldz2 : 0);
  p.t3 = take(p.layers3 ? p.kt * p.ldz2 : 0);
  p.w3 = take(p.layers3 ? 2 * h2 * h2 : 0);
  p.w3t = take(p.layers3 ? 2 * h2 * h2 : 0);
  p.b3 = take(p.layers3 ? h2 : 0);
  p.wg = take(p.wg_words);
  p.w1 = take(p.Kin1 * h1);
  p.w2 = take(p.Kin2 * h2);
  p.w2t = take(h2 * p.Kin2);
  p.b1 = take(h1);
  p.b2 = take(h2);
  p.fc2w = take(out * Hd);
  p.fc1b = take(Hd);
  p.fc2b = take(out);
  p.rrow = take(C2);
  p.hrow = take(Hd);
  p.dhrow = take(Hd);
  p.drrow = take(h2);
  p.prow = take(out);
  p.red = take((S3_THREADS / 32) * h2 > 8 ? (S3_THREADS / 32) * h2 : 8);
  p.rpart = take(h2);
  p.s1 = take(p.kt);
  p.post1 = take(p.kt);
  p.arg0 = take(p.kt * h1);
  p.arg1 = take(p.qt * h2);
  p.blob_words = stage ? DRGNN_BLOB_USED(max_n, max_e) : 0;
  p.blob = take(p.blob_words);
  p.wblob = take(kind == 1 ? p.blob_words : 0);
  p.bases = take(2 * 9 * S3_MAX_TILES);           // nine distributed arrays x 8 tiles of 8-byte generic pointers
  p.bars = take(4);
  p.total = o;
  p.fused_reduce = 0;
  return p;
}

// ---------------------------------------------------------------------------------------------------------
// A row-distributed array: tile t holds rows [t*rpt, (t+1)*rpt) at base[t] (a generic pointer into that CTA's
// shared memory, own or remote).  Level-0 features with NT > 1 are one global array: rpt = INT_MAX, base[0].
struct S3Rows {
  const void* b0;            // base of tile 0 (distributed: its address in the cluster's shared-memory window)
  long long stride;          // bytes from one tile's base to the next (the window is linear in the CTA rank); 0: one array
  unsigned magic;            // ceil(2^32 / rpt): owner(i) = umulhi(i, magic), exact for i, rpt < 2^16
  int rpt;
  int ld;
  __device__ __forceinline__ const void* row(int i) const {
    if (stride == 0) return reinterpret_cast<const char*>(b0) + (size_t)(i * ld) * 4u;
    const int o = rpt == 1 ? i : (int)__umulhi((unsigned)i, magic);    // (2^32 / 1 does not fit the magic word)
    return reinterpret_cast<const char*>(b0) + o * stride + (size_t)((i - o * rpt) * ld) * 4u;
  }
  __device__ __forceinline__ const float* frow(int i) const { return reinterpret_cast<const float*>(row(i)); }
  __device__ __forceinline__ const int* irow(int i) const { return reinterpret_cast<const int*>(row(i)); }
};
// magic word of a rows-per-tile count (one 64-bit division: computed ONCE per kernel for the three levels, not per
// call - the division was 5 % of the kernel's stall samples at cfg4)
__device__ __forceinline__ unsigned s3_magic(int tiles, int rpt) {
  return (tiles > 1 && rpt > 1) ? (unsigned)((0x100000000ull + (unsigned)rpt - 1ull) / (unsigned)rpt) : 0u;
}
// rows of a cluster-distributed array (tile t at base[t] = base[0] + t * stride, rpt rows each): the owner's base is
// ONE multiply-add (a pointer table in shared memory cost a dependent load per row access: 8 % of the stall samples
// at cfg4); one tile: direct addressing
__device__ __forceinline__ S3Rows s3_rows(const void* const* base, int tiles, int rpt, int ld, unsigned magic) {
  S3Rows r;
  r.b0 = base[0];
  r.stride = tiles > 1 ? (reinterpret_cast<const char*>(base[1]) - reinterpret_cast<const char*>(base[0])) : 0;
  r.rpt = rpt;
  r.ld = ld;
  r.magic = magic;
  return r;
}
__device__ __forceinline__ S3Rows s3_rows_flat(const void* b0, int ld) {
  S3Rows r;
  r.b0 = b0;
  r.stride = 0;
  r.rpt = 0;
  r.ld = ld;
  r.magic = 0u;
  return r;
}
// item -> (item / w, item % w) with a shift when w is a power of two (the usual widths: 4, 8, 16)
struct S3Div {
  int w, sh;
  __device__ __forceinline__ explicit S3Div(int w_) : w(w_), sh(-1) {
    if ((w_ & (w_ - 1)) == 0) sh = __ffs(w_) - 1;
  }
  __device__ __forceinline__ void split(int item, int& q, int& r) const {
    if (sh >= 0) { q = item >> sh; r = item & (w - 1); }
    else { q = item / w; r = item - q * w; }
  }
};

// Aggregation of one conv layer over the CTA's rows [lo, hi): writes zin (local row i - lo) as documented at
// the top, and (s_out / post_out != NULL) the per-row scalars the backward of conv2 needs.  W4 = C / 4 lanes
// per row, each owning 4 channels; every lane walks the row's CSR slice (ascending slot = the CPU scatter order).
__device__ __noinline__ void s3_aggregate(int kind, const int* __restrict__ rp, const int* __restrict__ col,
                                          const float* __restrict__ ew, S3Rows src, int C, int lo, int hi,
                                          float* __restrict__ zin, int ldz, float* __restrict__ s_out,
                                          float* __restrict__ post_out, int tid, int nth) {
  const int W4 = C >> 2;
  const int rows = hi - lo;
  const S3Div dv(W4);
#pragma unroll 1
  for (int item = tid; item < rows * W4; item += nth) {
    int il, q4;
    dv.split(item, il, q4);
    const int i = lo + il;
    const int sb = rp[i], se = rp[i + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float wsum = 0.f;
    if (kind == 1) {
#pragma unroll 4
      for (int p = sb; p < se; ++p) {
        const float w = ew[p];
        const float4 v = *reinterpret_cast<const float4*>(src.frow(col[p]) + q4 * 4);
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
        wsum += w;
      }
    } else {
#pragma unroll 4
      for (int p = sb; p < se; ++p) {
        const float4 v = *reinterpret_cast<const float4*>(src.frow(col[p]) + q4 * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    float* zr = zin + il * ldz;
    if (kind == 0) {
      *reinterpret_cast<float4*>(zr + q4 * 4) = acc;
    } else {
      const int deg = se - sb;
      // post: 1/max(deg,1) (scatter_mean, sGAT.py:81) | 1/deg with deg = 0 -> inf, 0 * inf = NaN (foutnet.py:73)
      const float post = kind == 1 ? 1.f / (float)max(deg, 1) : 1.f / (float)deg;
      const float selfc = kind == 1 ? post * wsum : 1.f;
      acc.x *= post; acc.y *= post; acc.z *= post; acc.w *= post;
      const float4 sv = *reinterpret_cast<const float4*>(src.frow(i) + q4 * 4);
      *reinterpret_cast<float4*>(zr + q4 * 4) = make_float4(selfc * sv.x, selfc * sv.y, selfc * sv.z, selfc * sv.w);
      *reinterpret_cast<float4*>(zr + C + q4 * 4) = acc;
      if (q4 == 0) {
        *reinterpret_cast<float4*>(zr + 2 * C) = make_float4(1.f, 0.f, 0.f, 0.f);   // ones column: bias gradient
        if (s_out) s_out[il] = selfc;
        if (post_out) post_out[il] = post;
      }
    }
  }
}

// C[m][n..n+3] = act( sum_k A[m*lda + k] * Bm[k*ldb + n] + bias[n] ) (* rscale[m] for columns >= scol)
// 2 x 4 register tile (same fmaf chain per output, ascending k, as step2 / the op-level linear kernel).
__device__ __noinline__ void s3_gemm(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb, int M, int N, int K,
                                     float* __restrict__ C, int ldc, const float* __restrict__ bias, int relu,
                                     const float* __restrict__ rscale, int scol, int tid, int nth) {
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
    if (bias) {
      const float4 bv = *reinterpret_cast<const float4*>(bias + ng * 4);
      acc0.x += bv.x; acc0.y += bv.y; acc0.z += bv.z; acc0.w += bv.w;
      acc1.x += bv.x; acc1.y += bv.y; acc1.z += bv.z; acc1.w += bv.w;
    }
    if (relu) {   // v < 0 ? 0 : v keeps NaN like torch.relu (Fout rows without neighbour)
      acc0.x = acc0.x < 0.f ? 0.f : acc0.x; acc0.y = acc0.y < 0.f ? 0.f : acc0.y; acc0.z = acc0.z < 0.f ? 0.f : acc0.z; acc0.w = acc0.w < 0.f ? 0.f : acc0.w;
      acc1.x = acc1.x < 0.f ? 0.f : acc1.x; acc1.y = acc1.y < 0.f ? 0.f : acc1.y; acc1.z = acc1.z < 0.f ? 0.f : acc1.z; acc1.w = acc1.w < 0.f ? 0.f : acc1.w;
    }
    const int m = mg * 2;
    if (rscale && ng * 4 >= scol) {
      const float r0 = rscale[m], r1 = (m + 1 < M) ? rscale[m + 1] : 0.f;
      acc0.x *= r0; acc0.y *= r0; acc0.z *= r0; acc0.w *= r0;
      acc1.x *= r1; acc1.y *= r1; acc1.z *= r1; acc1.w *= r1;
    }
    *reinterpret_cast<float4*>(C + m * ldc + ng * 4) = acc0;
    if (m + 1 < M) *reinterpret_cast<float4*>(C + (m + 1) * ldc + ng * 4) = acc1;
  }
}

// Cluster max + argmax over the CTA's clusters [lo, hi): members anywhere in the graph (row-distributed src).
// First member wins ties, a NaN never wins, empty -> 0 (torch_scatter's CPU scatter_max).
__device__ __noinline__ void s3_cluster_max(const int* __restrict__ cmp, const int* __restrict__ cmem, S3Rows src, int lo, int hi,
                                            float* __restrict__ dst, int ldd, int* __restrict__ arg, int ldarg, int W4,
                                            int tid, int nth) {
  const int rows = hi - lo;
  const S3Div dv(W4);
#pragma unroll 1
  for (int item = tid; item < rows * W4; item += nth) {
    int kl, q4;
    dv.split(item, kl, q4);
    const int k = lo + kl;
    const int sb = cmp[k], se = cmp[k + 1];
    float4 best = make_float4(-FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX);
    int4 am = make_int4(-1, -1, -1, -1);
#pragma unroll 1
    for (int p = sb; p < se; ++p) {
      const int i = cmem[p];
      const float4 v = *reinterpret_cast<const float4*>(src.frow(i) + q4 * 4);
      if (v.x > best.x) { best.x = v.x; am.x = i; }
      if (v.y > best.y) { best.y = v.y; am.y = i; }
      if (v.z > best.z) { best.z = v.z; am.z = i; }
      if (v.w > best.w) { best.w = v.w; am.w = i; }
    }
    if (am.x < 0) best.x = 0.f;
    if (am.y < 0) best.y = 0.f;
    if (am.z < 0) best.z = 0.f;
    if (am.w < 0) best.w = 0.f;
    *reinterpret_cast<float4*>(dst + kl * ldd + q4 * 4) = best;
    *reinterpret_cast<int4*>(arg + kl * ldarg + q4 * 4) = am;
  }
}

// Backward of cluster max + ReLU, IN PLACE on the CTA's rows [lo, hi) of z:
//   z[i][c] <- (arg[cl[i]][c] == i && z[i][c] > 0) ? d[cl[i]][c] * scale : 0
// arg (and d, unless drow != NULL: one gradient row for every cluster, the read-out mean) are row-distributed.
__device__ __noinline__ void s3_route(const int* __restrict__ cl, S3Rows arg, S3Rows d, const float* __restrict__ drow, float scale,
                                      float* __restrict__ z, int ldz, int lo, int hi, int W4, int tid, int nth) {
  const int rows = hi - lo;
  const S3Div dv(W4);
#pragma unroll 1
  for (int item = tid; item < rows * W4; item += nth) {
    int il, q4;
    dv.split(item, il, q4);
    const int i = lo + il;
    const int k = cl[i];
    const int4 am = *reinterpret_cast<const int4*>(arg.irow(k) + q4 * 4);
    float* zp = z + il * ldz + q4 * 4;
    const float4 zz = *reinterpret_cast<const float4*>(zp);
    const float4 dd = drow ? *reinterpret_cast<const float4*>(drow + q4 * 4) : *reinterpret_cast<const float4*>(d.frow(k) + q4 * 4);
    float4 v;
    v.x = (am.x == i && zz.x > 0.f) ? dd.x * scale : 0.f;
    v.y = (am.y == i && zz.y > 0.f) ? dd.y * scale : 0.f;
    v.z = (am.z == i && zz.z > 0.f) ? dd.z * scale : 0.f;
    v.w = (am.w == i && zz.w > 0.f) ? dd.w * scale : 0.f;
    *reinterpret_cast<float4*>(zp) = v;
  }
}

// Transposed aggregation of conv2's backward over the CTA's coarsened rows [lo, hi):
//   dp1[j] = selfc_j * dzin2[j][0:C]  (kinds 1, 2)  +  sum_{p in csc(j)} w[p] * dzin2[row_p][goff : goff + C]
// (the aggregated half of dzin2 is already multiplied by post[row], see s3_gemm's rscale).
__device__ __noinline__ void s3_gather_t(int kind, const int* __restrict__ cp, const int* __restrict__ crow,
                                         const float* __restrict__ ew, S3Rows src, int goff, const float* __restrict__ selfloc,
                                         int ldself, const float* __restrict__ s1, int C, int lo, int hi, float* __restrict__ dst,
                                         int ldd, int tid, int nth) {
  const int W4 = C >> 2;
  const int rows = hi - lo;
  const S3Div dv(W4);
#pragma unroll 1
  for (int item = tid; item < rows * W4; item += nth) {
    int jl, q4;
    dv.split(item, jl, q4);
    const int j = lo + jl;
    const int sb = cp[j], se = cp[j + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kind == 1) {
#pragma unroll 4
      for (int p = sb; p < se; ++p) {
        const float w = ew[p];
        const float4 v = *reinterpret_cast<const float4*>(src.frow(crow[p]) + goff + q4 * 4);
        acc.x = fmaf(w, v.x, acc.x); acc.y = fmaf(w, v.y, acc.y); acc.z = fmaf(w, v.z, acc.z); acc.w = fmaf(w, v.w, acc.w);
      }
    } else {
#pragma unroll 4
      for (int p = sb; p < se; ++p) {
        const float4 v = *reinterpret_cast<const float4*>(src.frow(crow[p]) + goff + q4 * 4);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
    }
    if (kind != 0) {
      const float sc = kind == 1 ? s1[jl] : 1.f;
      const float4 sv = *reinterpret_cast<const float4*>(selfloc + jl * ldself + q4 * 4);
      acc.x = fmaf(sc, sv.x, acc.x); acc.y = fmaf(sc, sv.y, acc.y); acc.z = fmaf(sc, sv.z, acc.z); acc.w = fmaf(sc, sv.w, acc.w);
    }
    *reinterpret_cast<float4*>(dst + jl * ldd + q4 * 4) = acc;
  }
}

// Split-K partial products of C[m][n] = sum_{k<K} At[k*lda + m] * Bm[k*ldb + n]  (M % 4 == 0, N % 4 == 0).
__device__ __noinline__ void s3_splitk_partial(const float* __restrict__ At, int lda, const float* __restrict__ Bm, int ldb, int M,
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
__device__ __noinline__ void s3_splitk_reduce(const float* __restrict__ scratch, int MN, int KS, float* __restrict__ dst, int tid,
                                              int nth) {
#pragma unroll 1
  for (int e = tid; e < MN; e += nth) {
    float acc = 0.f;
#pragma unroll 4
    for (int s = 0; s < KS; ++s) acc += scratch[(size_t)s * MN + e];
    dst[e] = acc;
  }
}

// Sum of the tiles' local weight-gradient matrices wg[M][N] (tile order) for this CTA's slice of the elements,
// stored to the graph's gradient row: rows m < Mw -> dw[m*N + n], row Mw -> db[n] (the ones-column row), rest dropped.
__device__ __noinline__ void s3_cross_tile_store(const void* const* wgbase, int tiles, int ti, int M, int N, int Mw,
                                                 float* __restrict__ dw, float* __restrict__ db, int tid, int nth) {
  const int MN = M * N;
  const int per = (MN + tiles - 1) / tiles;
  const int e0 = ti * per, e1 = min(MN, e0 + per);
  const int wend = Mw * N, bend = wend + N;     // [0, wend): weight rows, [wend, bend): the bias row
#pragma unroll 1
  for (int e = e0 + tid; e < e1; e += nth) {
    float acc = 0.f;
#pragma unroll 1
    for (int tt = 0; tt < tiles; ++tt) acc += reinterpret_cast<const float*>(wgbase[tt])[e];
    if (e < wend) dw[e] = acc;
    else if (e < bend && db) db[e - wend] = acc;
  }
}

// rows x W4 16-byte words of shared memory -> global memory (test mirror of the intermediates).
// iadd1 > 0: the words are int32 local ids; non-negative ones are stored as id + (iadd1 - 1).
__device__ __noinline__ void s3_mirror(const void* src, int lds, void* dst, int64_t ldd, int rows, int W4, int iadd1, int tid,
                                       int nth) {
  const int* sp = reinterpret_cast<const int*>(src);
  int* dp = reinterpret_cast<int*>(dst);
#pragma unroll 1
  for (int item = tid; item < rows * W4; item += nth) {
    const int i = item / W4, q4 = item - i * W4;
    int4 v = *reinterpret_cast<const int4*>(sp + i * lds + q4 * 4);
    if (iadd1 > 0) {
      const int ad = iadd1 - 1;
      v.x = v.x >= 0 ? v.x + ad : v.x; v.y = v.y >= 0 ? v.y + ad : v.y;
      v.z = v.z >= 0 ? v.z + ad : v.z; v.w = v.w >= 0 ? v.w + ad : v.w;
    }
    *reinterpret_cast<int4*>(dp + i * ldd + q4 * 4) = v;
  }
}

// Gradient reduction (+ Adam, + peer exchange) behind a grid barrier: the code of step2's tail, for any grid
// whose CTAs are all co-resident.  psum: [4][128] floats of scratch, red: >= 8 floats.
__device__ __noinline__ void s3_grid_reduce(const drgnn_net_step_args& s, const drgnn_peer_comm& C, float* __restrict__ psum,
                                            float* __restrict__ red) {
  const int t = threadIdx.x;
  constexpr int T = S3_THREADS;
  __syncthreads();
  unsigned* sync_ctr = reinterpret_cast<unsigned*>(s.step_dev + 2);
  if (t == 0) {
    __threadfence();
    atomicAdd(sync_ctr, 1u);
    const unsigned G = gridDim.x;
    const unsigned long long t0 = s3_globaltimer();
    while (s3_ld_acquire(sync_ctr) < G) {
      if (s3_globaltimer() - t0 > 2000000000ull) {
        atomicOr(s.status, 128);
        break;
      }
    }
  } else if (t == 32) {
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
  const int n = s.n_params, B = s.B;
  const int per = (n + 1 + (int)gridDim.x - 1) / (int)gridDim.x;
  const int el = t & 127, q = t >> 7;
  float* adamc = red;
  unsigned* epoch_s = reinterpret_cast<unsigned*>(red + 4);
  const int world = C.world, rank = C.rank;
  const bool peers = world > 1;
  if (t == T - 1) {
    unsigned* ticket = reinterpret_cast<unsigned*>(s.step_dev + 1);
    if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
      __threadfence();
      *ticket = 0u;
      *sync_ctr = 0u;
      if (peers) *reinterpret_cast<volatile uint32_t*>(C.ctr) = *epoch_s;
      if (s.fuse_adam) s.step_dev[0] = adamc[0];
    }
  }
#pragma unroll 1
  for (int sweep = 0; sweep < per; sweep += 128) {
    const int e = (int)blockIdx.x * per + sweep + el;
    const bool mine = sweep + el < per && e <= n;
    float acc = 0.f;
    float adam_mi = 0.f, adam_vi = 0.f, adam_pi = 0.f;
    if (q == 0 && mine && !peers && s.fuse_adam && e < n) {
      adam_mi = __ldcg(s.adam_m + e); adam_vi = __ldcg(s.adam_v + e); adam_pi = __ldcg(s.adam_p + e);
    }
    if (mine) {
      const int gs = (B + 3) >> 2;
      const int g0 = q * gs, g1 = min(B, g0 + gs);
      const float* src = s.partial + e;
      int gg = g0;
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
        const unsigned ep = *epoch_s;
        const int64_t slot = ((int64_t)(ep & 1u) * world + rank) * C.stride + e;
#pragma unroll 1
        for (int p = 0; p < world; ++p) s3_st_ll(C.xll[(rank + p) % world] + slot, __float_as_uint(acc), ep);
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
    const unsigned epoch = *epoch_s;
    const int par = (int)(epoch & 1u);
    const unsigned long long t0 = s3_globaltimer(), limit = C.timeout_ns ? C.timeout_ns : 20000000000ull;
#pragma unroll 1
    for (int sweep = 0; sweep < per; sweep += T) {
      const int e = (int)blockIdx.x * per + sweep + t;
      if (sweep + t < per && e <= n) {
        float tot = 0.f;
        const uint64_t* mine = C.xll[rank] + (int64_t)par * world * C.stride + e;
#pragma unroll 1
        for (int rr = 0; rr < world; ++rr) {
          unsigned bits, ep;
          s3_ld_ll(mine + (int64_t)rr * C.stride, bits, ep);
          while (ep != epoch) {
            if (s3_globaltimer() - t0 > limit) {
              atomicOr(C.ctr + 2, 1u);
              break;
            }
            __nanosleep(20);
            s3_ld_ll(mine + (int64_t)rr * C.stride, bits, ep);
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

__host__ __device__ inline int s3_split(int cap_words, int mn) {
  int ks = cap_words / mn;
  if (ks > S3_MAX_KS) ks = S3_MAX_KS;
  if (ks < 1) ks = 1;
  return ks;
}

__global__ void __launch_bounds__(S3_THREADS, 1)
    net_graph_step3_kernel(const drgnn_net_step_args s, const Step3Plan P, const drgnn_peer_comm C) {
  extern __shared__ __align__(16) float sm[];
  cgx::cluster_group cluster = cgx::this_cluster();
  const int kind = s.kind;
  const int NT = P.tiles, CS = P.cs;
  const int r = (int)cluster.block_rank();
  const int br = r / NT, ti = r - br * NT;      // branch (GINet), node tile
  int g = blockIdx.x / CS;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if ((s.flags & 32) && s.B + 4 <= P.scr_words) {
    // Largest graph first (flag bit 5; grids larger than the device): cluster c takes the graph of size rank c, so
    // the clusters scheduled last - when SMs free up - run the smallest graphs and the tail of the launch is short
    // (longest-processing-time order; mixed sizes, BASELINE config 5).  Every CTA ranks the B node counts itself
    // (one coalesced load of node_ptr, B^2 / T shared-memory compares): no extra launch, no host-side permutation.
    int* sz = reinterpret_cast<int*>(sm + P.scr);
    const int B = s.B, cid = g;
#pragma unroll 1
    for (int i = t; i < B; i += S3_THREADS) sz[i] = __ldg(s.node_ptr + i + 1) - __ldg(s.node_ptr + i);
    __syncthreads();
#pragma unroll 1
    for (int i = t; i < B; i += S3_THREADS) {
      const int ni = sz[i];
      int rank = 0;
#pragma unroll 4
      for (int h = 0; h < B; ++h) {
        const int nh = sz[h];
        rank += (nh > ni || (nh == ni && h < i)) ? 1 : 0;
      }
      if (rank == cid) sz[B] = i;
    }
    __syncthreads();
    g = sz[B];
    __syncthreads();
  }
  const bool timers = (s.flags & 8) != 0 && blockIdx.x == 0 && t == 0;   // phase clocks of block 0 (diagnostic)
  const bool cta_times = (s.flags & 8) != 0 && t == 0 && blockIdx.x < 2048;
  if (cta_times) g_cta_times[blockIdx.x][0] = s3_globaltimer();
  constexpr int T = S3_THREADS, NW = S3_THREADS / 32;
  const int F = s.F, H1 = s.h1, H2 = s.h2, Hd = s.Hd, out = s.out;
  const int NBR = P.nbr, C1 = NBR * H1, C2 = NBR * H2;
  const int Kin1 = P.Kin1, Kin2 = P.Kin2;
  const int co1 = br * H1, co2 = br * H2;
  const bool mirror = (s.flags & 1) != 0;
  const bool multi = NT > 1;             // rows distributed over the cluster: cluster barriers between phases
  const bool tc = (s.flags & 4) != 0;    // dense products on the tensor cores (mma.sync 3xTF32) instead of FFMA tiles
  const bool staged = P.blob_words > 0;  // the graph's blob and feature tile are staged in shared memory
  DRGNN_PHASE3(0);
  float* xs = sm + P.xs;     float* scr = sm + P.scr;   float* zin1 = sm + P.zin1; float* z1 = sm + P.z1;
  float* p1 = sm + P.p1;     float* zin2 = sm + P.zin2; float* z2 = sm + P.z2;     float* p2 = sm + P.p2;
  float* wg = sm + P.wg;     float* w1 = sm + P.w1;     float* w2 = sm + P.w2;     float* w2t = sm + P.w2t;
  float* b1 = sm + P.b1;     float* b2 = sm + P.b2;     float* fc2w = sm + P.fc2w; float* fc1b = sm + P.fc1b;
  float* fc2b = sm + P.fc2b; float* rrow = sm + P.rrow; float* hrow = sm + P.hrow; float* dhrow = sm + P.dhrow;
  float* drrow = sm + P.drrow; float* prow = sm + P.prow; float* red = sm + P.red; float* rpart = sm + P.rpart;
  float* s1 = sm + P.s1;     float* post1 = sm + P.post1;
  int* ism = reinterpret_cast<int*>(sm);
  int* arg0 = ism + P.arg0;  int* arg1 = ism + P.arg1;
  int* blbs = ism + P.blob;  float* wbls = sm + P.wblob;
  const void** bases = reinterpret_cast<const void**>(ism + P.bases);   // [7][S3_MAX_TILES]
  uint64_t* bars = reinterpret_cast<uint64_t*>(ism + P.bars);
  const int LDZIN1 = P.ldzin1, LDZ1 = P.ldz1, LDP = P.ldp, LDZIN2 = P.ldzin2, LDZ2 = P.ldz2;
  const bool L3 = P.layers3 != 0;
  float* zin3 = sm + P.zin3; float* z3 = sm + P.z3; float* t3 = sm + P.t3; float* w3 = sm + P.w3; float* w3t = sm + P.w3t;
  float* b3 = sm + P.b3;
  const int LDZIN3 = P.ldzin3;
  float* dp1 = p1;          // the pooled features are dead once conv2 has aggregated them
  float* dzin2 = zin2;      // overwritten after the conv2 weight-gradient products

  // ---- this CTA's weights in the layouts the products read (B operand: [k][n]): 4-byte asynchronous copies whose
  // DESTINATION address does the transposition - no register, no dependent store, so every copy of the phase is in
  // flight at once (ONE L2 round trip instead of one per loop), waited for together with the bulk copies
  {
    const float* W1g = s.params + s.off_w1;
    const float* W2g = s.params + s.off_w2;
    if (kind == 0) {
#pragma unroll 1
      for (int i = t; i < H1 * F; i += T) {        // W1 [2][H1][F] branch br (contiguous) -> w1 [F][H1]
        const int c = i / F, f = i - c * F;
        s3_cp4(w1 + f * H1 + c, W1g + (int64_t)co1 * F + i);
      }
#pragma unroll 1
      for (int i = t; i < H2 * H1; i += T) {       // W2 [2][H2][H1] branch br -> w2t [H2][H1] (as stored), w2 [H1][H2]
        const int o = i / H1, j = i - o * H1;
        const float* src = W2g + (int64_t)br * H2 * H1 + i;
        s3_cp4(w2t + i, src);
        s3_cp4(w2 + j * H2 + o, src);
      }
    } else {
#pragma unroll 1
      for (int i = t; i < Kin1 * H1; i += T) s3_cp4(w1 + i, W1g + i);     // [2F][H1] as stored
#pragma unroll 1
      for (int i = t; i < Kin2 * H2; i += T) {     // [2H1][H2] as stored -> w2, transposed -> w2t [H2][2H1]
        const int k = i / H2, o = i - k * H2;
        s3_cp4(w2 + i, W2g + i);
        s3_cp4(w2t + o * Kin2 + k, W2g + i);
      }
      if (L3) {
        const float* W3g = s.params + s.off_w3;
#pragma unroll 1
        for (int i = t; i < 2 * H2 * H2; i += T) {   // [2H2][H2] as stored -> w3, transposed -> w3t [H2][2H2]
          const int k = i / H2, o = i - k * H2;
          s3_cp4(w3 + i, W3g + i);
          s3_cp4(w3t + o * 2 * H2 + k, W3g + i);
        }
#pragma unroll 1
        for (int i = t; i < H2; i += T) s3_cp4(b3 + i, s.params + s.off_b3 + i);
      }
#pragma unroll 1
      for (int i = t; i < H1; i += T) s3_cp4(b1 + i, s.params + s.off_b1 + i);
#pragma unroll 1
      for (int i = t; i < H2; i += T) s3_cp4(b2 + i, s.params + s.off_b2 + i);
    }
#pragma unroll 1
    for (int i = t; i < out * Hd; i += T) s3_cp4(fc2w + i, s.params + s.off_fc2w + i);
#pragma unroll 1
    for (int i = t; i < Hd; i += T) s3_cp4(fc1b + i, s.params + s.off_fc1b + i);
#pragma unroll 1
    for (int i = t; i < out; i += T) s3_cp4(fc2b + i, s.params + s.off_fc2b + i);
  }
  // ---- graph extents
  int n0, n, eg0, m;
  if (s.gdesc) {
    const int4 lo = __ldg(reinterpret_cast<const int4*>(s.gdesc + 8 * (int64_t)g));
    const int4 hi = __ldg(reinterpret_cast<const int4*>(s.gdesc + 8 * (int64_t)g + 4));
    n0 = lo.w; eg0 = hi.x; m = hi.y; n = hi.w;
  } else {
    n0 = __ldg(s.node_ptr + g); n = __ldg(s.node_ptr + g + 1) - n0;
    eg0 = __ldg(s.edge_ptr + g); m = __ldg(s.edge_ptr + g + 1) - eg0;
  }
  const bool train = !(s.forward_only || s.task == 0);
  float* part = s.partial + (int64_t)g * s.partial_ld;
  const uint32_t drop_ctr = (!s.keep && s.drop_p > 0.f) ? (uint32_t)__ldg(s.step_dev) : 0u;
  bool valid = true;

  do {
  if (n < 0 || m < 0 || n > s.max_n || m > s.max_e) {   // host bounds violated: flag, contribute nothing
    asm volatile("cp.async.wait_all;" ::: "memory");      // the weight copies issued above
    if (t == 0) atomicOr(s.status, 64);
    valid = false;
    break;
  }
  const int64_t boff = DRGNN_BLOB_OFFSET(g, n0, eg0);
  const int* blb = s.blob + boff;                 // NT > 1: the index lists are read from global memory / L2
  const float* wbl = s.wblob ? s.wblob + boff : nullptr;
  const bool pre = s.zin1 != nullptr;             // conv1's input rows were computed by the structure pass
  if (t == 0 && (staged || pre)) {
    if (staged) s3_mbar_init(&bars[0], 1);
    if (pre) s3_mbar_init(&bars[1], 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the initialised barriers, seen by the async proxy
    if (staged) {
      const uint32_t bbytes = (uint32_t)DRGNN_BLOB_USED(n, m) * 4u, xbytes = pre ? 0u : (uint32_t)(n * F) * 4u;
      s3_mbar_expect_tx(&bars[0], bbytes + xbytes + (kind == 1 ? bbytes : 0u));
      s3_bulk_g2s(blbs, blb, bbytes, &bars[0]);
      if (kind == 1) s3_bulk_g2s(wbls, wbl, bbytes, &bars[0]);
      if (xbytes) s3_bulk_g2s(xs, s.x + (int64_t)n0 * F, xbytes, &bars[0]);
    }
    if (pre) {   // this tile's rows of zin1 (row stride LDZIN1), ONE bulk copy
      const int nta_ = s3_cdiv(max(n, 1), NT);
      const int lo_ = min(n, ti * nta_), hi_ = min(n, lo_ + nta_);
      const uint32_t zbytes = (uint32_t)((hi_ - lo_) * LDZIN1) * 4u;
      s3_mbar_expect_tx(&bars[1], zbytes);
      if (zbytes) s3_bulk_g2s(zin1, s.zin1 + ((int64_t)n0 + lo_) * LDZIN1, zbytes, &bars[1]);
    }
  }
  if (staged) {
    blb = blbs;
    wbl = wbls;
  }
  // ---- DSMEM base pointers of the row-distributed arrays of this branch: z1, p1, arg0, z2, arg1, dzin2/zin2, wg
  if (t < 9 * NT) {
    const int which = t / NT, tt = t - which * NT;
    float* local = which == 0 ? z1 : which == 1 ? p1 : which == 2 ? reinterpret_cast<float*>(arg0)
                 : which == 3 ? z2 : which == 4 ? reinterpret_cast<float*>(arg1) : which == 5 ? zin2
                 : which == 6 ? wg : which == 7 ? z3 : zin3;
    // (every tile through the cluster window, the own one included: base[t] = base[0] + t * stride, see S3Rows)
    bases[which * S3_MAX_TILES + tt] = NT == 1 ? local : cluster.map_shared_rank(local, (unsigned)(br * NT + tt));
  }
  asm volatile("cp.async.wait_all;" ::: "memory");   // this thread's weight copies
  __syncthreads();
  if (NT > 2 && t < 9 * NT) {   // the window must be linear in the CTA rank (it is: rank in the upper address bits)
    const int which = t / NT, tt = t - which * NT;
    const char* b0_ = reinterpret_cast<const char*>(bases[which * S3_MAX_TILES]);
    const char* b1_ = reinterpret_cast<const char*>(bases[which * S3_MAX_TILES + 1]);
    if (reinterpret_cast<const char*>(bases[which * S3_MAX_TILES + tt]) != b0_ + tt * (b1_ - b0_)) atomicOr(s.status, 64);
  }
  if (staged) s3_mbar_wait(&bars[0], 0);
  if (pre) s3_mbar_wait(&bars[1], 0);
  const int K = blb[2], E1 = blb[3], Q = blb[4];
  if (blb[5] != 1 || blb[0] != n || blb[1] != m || K > s.max_k || Q > s.max_q || K < 0 || Q < 0 || E1 < 0 || E1 > m) {
    if (t == 0) atomicOr(s.status, 64);
    valid = false;
    break;
  }
  const BlobLayout BL = blob_layout(n, m);
  const int* rp0 = blb + BL.rp0;     const int* col0 = blb + BL.col0;   const int* rp1 = blb + BL.rp1;  const int* col1 = blb + BL.col1;
  const int* cmp0 = blb + BL.cmp0;   const int* cmem0 = blb + BL.cmem0; const int* cl0 = blb + BL.cl0;
  const int* cmp1 = blb + BL.cmp1;   const int* cmem1 = blb + BL.cmem1; const int* cl1 = blb + BL.cl1;
  const int* cscp1 = blb + BL.cscp1; const int* cscr1 = blb + BL.cscr1;
  const float* ew0 = wbl ? wbl + BL.col0 : nullptr;
  const float* ew1 = wbl ? wbl + BL.col1 : nullptr;
  const float* ew1t = wbl ? wbl + BL.cscr1 : nullptr;
  // ---- this tile's rows at the three levels
  const int nta = s3_cdiv(max(n, 1), NT), kta = s3_cdiv(max(K, 1), NT), qta = s3_cdiv(max(Q, 1), NT);
  const int lo0 = min(n, ti * nta), hi0 = min(n, lo0 + nta);
  const int lo1 = min(K, ti * kta), hi1 = min(K, lo1 + kta);
  const int lo2 = min(Q, ti * qta), hi2 = min(Q, lo2 + qta);
  const int r0n = hi0 - lo0, r1n = hi1 - lo1, r2n = hi2 - lo2;
  const unsigned mg0 = s3_magic(NT, nta), mg1 = s3_magic(NT, kta), mg2 = s3_magic(NT, qta);
  const void* const* bz1 = bases;                    const void* const* bp1 = bases + S3_MAX_TILES;
  const void* const* barg0 = bases + 2 * S3_MAX_TILES; const void* const* bz2 = bases + 3 * S3_MAX_TILES;
  const void* const* barg1 = bases + 4 * S3_MAX_TILES; const void* const* bdzin2 = bases + 5 * S3_MAX_TILES;
  const void* const* bwg = bases + 6 * S3_MAX_TILES;
  const void* const* bz3 = bases + 7 * S3_MAX_TILES; const void* const* bdzin3 = bases + 8 * S3_MAX_TILES;
  // level-0 features: the staged tile or the global rows of the graph
  const void* xsrc = staged ? (const void*)xs : (const void*)(s.x + (int64_t)n0 * F);
  if (multi) cluster.sync();   // every CTA of the cluster runs (its shared memory may be read from now on)
  DRGNN_PHASE3(1);
  const int F4 = F >> 2, H14 = H1 >> 2, H24 = H2 >> 2;

  // ---- conv1: aggregate, transform
  if (!pre) {
    s3_aggregate(kind, rp0, col0, ew0, s3_rows_flat(xsrc, F), F, lo0, hi0, zin1, LDZIN1, nullptr, nullptr, t, T);
    __syncthreads();
  }
  DRGNN_PHASE3(2);
  if (tc) tc_gemm(zin1, LDZIN1, w1, H1, r0n, H1, Kin1, z1, LDZ1, kind ? b1 : nullptr, 1, nullptr, 0, t, T);
  else s3_gemm(zin1, LDZIN1, w1, H1, r0n, H1, Kin1, z1, LDZ1, kind ? b1 : nullptr, 1, nullptr, 0, t, T);
  if (multi) cluster.sync(); else __syncthreads();
  DRGNN_PHASE3(3);
  // ---- P1 = cluster max of Z1 (community_pooling.py:201): members may live in any tile
  s3_cluster_max(cmp0, cmem0, s3_rows(bz1, NT, nta, LDZ1, mg0), lo1, hi1, p1, LDP, arg0, H1, H14, t, T);
  if (multi) cluster.sync(); else __syncthreads();
  DRGNN_PHASE3(4);
  // ---- conv2 on the coarsened graph
  s3_aggregate(kind, rp1, col1, ew1, s3_rows(bp1, NT, kta, LDP, mg1), H1, lo1, hi1, zin2, LDZIN2, s1, post1, t, T);
  __syncthreads();
  DRGNN_PHASE3(5);
  if (tc) tc_gemm(zin2, LDZIN2, w2, H2, r1n, H2, Kin2, z2, LDZ2, kind ? b2 : nullptr, 1, nullptr, 0, t, T);
  else s3_gemm(zin2, LDZIN2, w2, H2, r1n, H2, Kin2, z2, LDZ2, kind ? b2 : nullptr, 1, nullptr, 0, t, T);
  if (multi) cluster.sync(); else __syncthreads();
  DRGNN_PHASE3(6);
  if (L3) {   // third conv layer on the coarsened graph (BASELINE config 3: "sGAT 3-layer"), h2 -> h2
    s3_aggregate(kind, rp1, col1, ew1, s3_rows(bz2, NT, kta, LDZ2, mg1), H2, lo1, hi1, zin3, LDZIN3, nullptr, nullptr, t, T);
    __syncthreads();
    if (tc) tc_gemm(zin3, LDZIN3, w3, H2, r1n, H2, 2 * H2, z3, LDZ2, b3, 1, nullptr, 0, t, T);
    else s3_gemm(zin3, LDZIN3, w3, H2, r1n, H2, 2 * H2, z3, LDZ2, b3, 1, nullptr, 0, t, T);
    if (multi) cluster.sync(); else __syncthreads();
  }
  float* zl = L3 ? z3 : z2;                                  // the last conv output: pooled, read out
  // ---- P2 = level-1 cluster max (max_pool_x)
  s3_cluster_max(cmp1, cmem1, s3_rows(L3 ? bz3 : bz2, NT, kta, LDZ2, mg1), lo2, hi2, p2, H2, arg1, H2, H24, t, T);
  __syncthreads();
  DRGNN_PHASE3(7);
  if (mirror) {   // parity tests: the intermediates the op-level path leaves in global memory (global ids)
    const int k0 = __ldg(s.kptr0 + g), q0 = __ldg(s.kptr1 + g);
    if (br == 0) s3_mirror(zin1, LDZIN1, s.Zin1 + (int64_t)(n0 + lo0) * Kin1, Kin1, r0n, Kin1 >> 2, 0, t, T);
    s3_mirror(z1, LDZ1, s.Z1 + (int64_t)(n0 + lo0) * C1 + co1, C1, r0n, H14, 0, t, T);
    s3_mirror(arg0, H1, s.arg0 + (int64_t)(k0 + lo1) * C1 + co1, C1, r1n, H14, n0 + 1, t, T);
    if (kind == 0) s3_mirror(zin2, LDZIN2, s.Zin2 + (int64_t)(k0 + lo1) * C1 + co1, C1, r1n, H14, 0, t, T);
    else s3_mirror(zin2, LDZIN2, s.Zin2 + (int64_t)(k0 + lo1) * Kin2, Kin2, r1n, Kin2 >> 2, 0, t, T);
    s3_mirror(z2, LDZ2, s.Z2 + (int64_t)(k0 + lo1) * C2 + co2, C2, r1n, H24, 0, t, T);
    s3_mirror(arg1, H2, s.arg1 + (int64_t)(q0 + lo2) * C2 + co2, C2, r2n, H24, k0 + 1, t, T);
  }
  // ---- read-out: sum of this tile's level-1 clusters; the mean over all tiles and both branches follows
#pragma unroll 1
  for (int c = t; c < H2; c += T) {
    float acc = 0.f;
    for (int q = 0; q < r2n; ++q) acc += p2[q * H2 + c];
    rpart[c] = acc;
  }
  if (CS > 1) cluster.sync(); else __syncthreads();
  {
    const float invq = 1.f / (float)max(Q, 1);
#pragma unroll 1
    for (int c = t; c < C2; c += T) {
      const int b = c / H2, cc = c - b * H2;
      float acc = 0.f;
#pragma unroll 1
      for (int tt = 0; tt < NT; ++tt) {
        const int rk = b * NT + tt;
        const float* rp_ = (rk == r) ? rpart : cluster.map_shared_rank(rpart, (unsigned)rk);
        acc += rp_[cc];
      }
      acc *= invq;
      rrow[c] = acc;
      if (r == 0 && s.R) s.R[(int64_t)g * C2 + c] = acc;
    }
  }
  __syncthreads();
  DRGNN_PHASE3(8);
  // ---- fc1 (every CTA, identical results): four lanes per hidden unit, fc1.weight read from global memory / L2
  {
    const float* fc1w = s.params + s.off_fc1w;
    const int sub = t & 3;
#pragma unroll 1
    for (int j = t >> 2; j < Hd; j += T >> 2) {
      const float* wrow = fc1w + (int64_t)j * C2;
      float av = 0.f;
#pragma unroll 2
      for (int c = sub * 4; c < C2; c += 16) {
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wrow + c));
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
          v = s3_hash_uniform(s.seed, drop_ctr, (uint32_t)(g * Hd + j)) >= s.drop_p ? v * s.keep_scale : 0.f;
        }
        hrow[j] = v;
      }
    }
  }
  __syncthreads();
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
  DRGNN_PHASE3(9);
  if (!train) break;
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
      const int tc = (int)__ldg(s.y_class + g);
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
        const float d = p - s.y[(int64_t)g * out + c];
        lg += d * d;
        prow[c] = 2.f * d * s.inv_norm * dp;
      }
    }
    if (r == 0) part[s.n_params] = lg * s.inv_norm;
  }
  __syncthreads();
  // ---- head backward: dh (all units, every CTA); the gradient rows of hidden units [j0, j1) are written by CTA r
  const int jper = s3_cdiv(Hd, CS);
  const int j0 = min(Hd, r * jper), j1 = min(Hd, j0 + jper), nj = j1 - j0;
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
    const int o = i / max(nj, 1), j = j0 + (i - o * nj);
    part[s.off_fc2w + o * Hd + j] = prow[o] * hrow[j];
  }
  if (r == 0) {
#pragma unroll 1
    for (int o = t; o < out; o += T) part[s.off_fc2b + o] = prow[o];
  }
  __syncthreads();
  {   // fc1.weight gradient rows of this CTA's hidden units: dh[j] * R[g][:], 16-byte stores
    const int C24 = C2 >> 2;
#pragma unroll 1
    for (int i = t; i < nj * C24; i += T) {
      const int jj = i / C24, c4 = i - jj * C24;
      const float dh = dhrow[j0 + jj];
      float4 rv = *reinterpret_cast<const float4*>(rrow + c4 * 4);
      rv.x *= dh; rv.y *= dh; rv.z *= dh; rv.w *= dh;
      *reinterpret_cast<float4*>(part + s.off_fc1w + (j0 + jj) * C2 + c4 * 4) = rv;
    }
  }
  // dR[c] of this branch's channels = sum_j dh[j] fc1w[j][co2 + c]: warps split the hidden units, fixed-order sum
  {
    const float* fc1w = s.params + s.off_fc1w;
#pragma unroll 1
    for (int c = lane; c < H2; c += 32) {
      float acc = 0.f;
#pragma unroll 2
      for (int j = warp; j < Hd; j += NW) acc = fmaf(dhrow[j], __ldg(fc1w + (int64_t)j * C2 + co2 + c), acc);
      red[warp * H2 + c] = acc;
    }
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
  DRGNN_PHASE3(10);
  // ---- Fout rule: a row without neighbour has a NaN input row whose gradient is exactly 0 (ReLU mask and
  // max-pool never select a NaN): drop it from the weight-gradient products instead of 0 * NaN (linear.cu)
  if (kind == 2) {
#pragma unroll 1
    for (int item = t; item < r0n * F4; item += T) {
      const int il = item / F4, q4 = item - il * F4;
      if (rp0[lo0 + il + 1] == rp0[lo0 + il]) *reinterpret_cast<float4*>(zin1 + il * LDZIN1 + F + q4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll 1
    for (int item = t; item < r1n * H14; item += T) {
      const int il = item / H14, q4 = item - il * H14;
      if (rp1[lo1 + il + 1] == rp1[lo1 + il]) *reinterpret_cast<float4*>(zin2 + il * LDZIN2 + H1 + q4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    if (L3) {
#pragma unroll 1
      for (int item = t; item < r1n * H24; item += T) {
        const int il = item / H24, q4 = item - il * H24;
        if (rp1[lo1 + il + 1] == rp1[lo1 + il]) *reinterpret_cast<float4*>(zin3 + il * LDZIN3 + H2 + q4 * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
  }
  // ---- dZ2 (in place): read-out mean backward, routed to the arg-max member, gated by ReLU
  s3_route(cl1, s3_rows(barg1, NT, qta, H2, mg2), s3_rows(barg1, NT, qta, H2, mg2), drrow, 1.f / (float)max(Q, 1), zl, LDZ2, lo1, hi1, H24, t, T);
  __syncthreads();
  DRGNN_PHASE3(11);
  if (L3) {
    // ---- third layer backward: dW3 / db3, dzin3 = dZ3 W3^T (aggregated half times post[row]) over zin3, then
    // dZ2 = relu'(Z2) * (s1 dzin3[:, :H2] + A1^T-weighted dzin3[:, H2:])  (in place on z2)
    const int M3 = 2 * H2 + 4, N3 = H2;
    const int KS3 = s3_split(P.scr_words, M3 * N3);
    if (tc) tc_splitk_partial(zin3, LDZIN3, z3, LDZ2, M3, N3, r1n, KS3, scr, t, T);
    else s3_splitk_partial(zin3, LDZIN3, z3, LDZ2, M3, N3, r1n, KS3, scr, t, T);
    __syncthreads();
    s3_splitk_reduce(scr, M3 * N3, KS3, wg, t, T);
    if (tc) tc_gemm(z3, LDZ2, w3t, 2 * H2, r1n, 2 * H2, H2, zin3, LDZIN3, nullptr, 0, post1, H2, t, T);
    else s3_gemm(z3, LDZ2, w3t, 2 * H2, r1n, 2 * H2, H2, zin3, LDZIN3, nullptr, 0, post1, H2, t, T);
    if (multi) cluster.sync(); else __syncthreads();
    s3_cross_tile_store(bwg, NT, ti, M3, N3, 2 * H2, part + s.off_w3, part + s.off_b3, t, T);
    s3_gather_t(kind, cscp1, cscr1, ew1t, s3_rows(bdzin3, NT, kta, LDZIN3, mg1), H2, zin3, LDZIN3, s1, H2, lo1, hi1, t3, LDZ2, t, T);
    __syncthreads();
#pragma unroll 1
    for (int item = t; item < r1n * H24; item += T) {
      const int il = item / H24, q4 = item - il * H24;
      float4* zp = reinterpret_cast<float4*>(z2 + il * LDZ2 + q4 * 4);
      const float4 zz = *zp;
      const float4 dd = *reinterpret_cast<const float4*>(t3 + il * LDZ2 + q4 * 4);
      *zp = make_float4(zz.x > 0.f ? dd.x : 0.f, zz.y > 0.f ? dd.y : 0.f, zz.z > 0.f ? dd.z : 0.f, zz.w > 0.f ? dd.w : 0.f);
    }
    if (multi) cluster.sync(); else __syncthreads();   // nobody reads this tile's wg / dzin3 any more
  }
  // ---- conv2 weight (+ bias) gradient over this tile's rows: split-K partials, local sum, sum over the tiles
  const int M2 = kind == 0 ? H2 : Kin2 + 4, N2 = kind == 0 ? H1 : H2;
  const int M1 = kind == 0 ? H1 : Kin1 + 4, N1 = kind == 0 ? F : H1;
  const int KS2 = s3_split(P.scr_words, M2 * N2), KS1 = s3_split(P.scr_words, M1 * N1);
  if (tc) {
    if (kind == 0) tc_splitk_partial(z2, LDZ2, zin2, LDZIN2, M2, N2, r1n, KS2, scr, t, T);
    else tc_splitk_partial(zin2, LDZIN2, z2, LDZ2, M2, N2, r1n, KS2, scr, t, T);
  } else {
    if (kind == 0) s3_splitk_partial(z2, LDZ2, zin2, LDZIN2, M2, N2, r1n, KS2, scr, t, T);
    else s3_splitk_partial(zin2, LDZIN2, z2, LDZ2, M2, N2, r1n, KS2, scr, t, T);
  }
  __syncthreads();
  s3_splitk_reduce(scr, M2 * N2, KS2, wg, t, T);
  // ---- dzin2 = dZ2 W2 (GINet: [K][H1]) | dZ2 W^T ([K][2H1], aggregated half times post[row]) - over zin2
  if (tc) tc_gemm(z2, LDZ2, w2t, Kin2, r1n, Kin2, H2, dzin2, LDZIN2, nullptr, 0, kind ? post1 : nullptr, H1, t, T);
  else s3_gemm(z2, LDZ2, w2t, Kin2, r1n, Kin2, H2, dzin2, LDZIN2, nullptr, 0, kind ? post1 : nullptr, H1, t, T);
  if (multi) cluster.sync(); else __syncthreads();
  DRGNN_PHASE3(12);
  if (kind == 0) s3_cross_tile_store(bwg, NT, ti, M2, N2, M2, part + s.off_w2 + br * H2 * H1, nullptr, t, T);
  else s3_cross_tile_store(bwg, NT, ti, M2, N2, Kin2, part + s.off_w2, part + s.off_b2, t, T);
  // ---- dP1 = transposed aggregation of dzin2 (CSC of the coarsened graph) (+ self term)
  s3_gather_t(kind, cscp1, cscr1, ew1t, s3_rows(bdzin2, NT, kta, LDZIN2, mg1), kind ? H1 : 0, dzin2, LDZIN2, s1, H1, lo1, hi1, dp1, LDP, t, T);
  if (multi) cluster.sync(); else __syncthreads();
  DRGNN_PHASE3(13);
  // ---- dZ1 (in place): routed to the arg-max node of its cluster, gated by ReLU
  s3_route(cl0, s3_rows(barg0, NT, kta, H1, mg1), s3_rows(bp1, NT, kta, LDP, mg1), nullptr, 1.f, z1, LDZ1, lo0, hi0, H14, t, T);
  __syncthreads();
  DRGNN_PHASE3(14);
  // ---- conv1 weight (+ bias) gradient
  if (tc) {
    if (kind == 0) tc_splitk_partial(z1, LDZ1, zin1, LDZIN1, M1, N1, r0n, KS1, scr, t, T);
    else tc_splitk_partial(zin1, LDZIN1, z1, LDZ1, M1, N1, r0n, KS1, scr, t, T);
  } else {
    if (kind == 0) s3_splitk_partial(z1, LDZ1, zin1, LDZIN1, M1, N1, r0n, KS1, scr, t, T);
    else s3_splitk_partial(zin1, LDZIN1, z1, LDZ1, M1, N1, r0n, KS1, scr, t, T);
  }
  __syncthreads();
  s3_splitk_reduce(scr, M1 * N1, KS1, wg, t, T);
  if (multi) cluster.sync(); else __syncthreads();
  if (kind == 0) s3_cross_tile_store(bwg, NT, ti, M1, N1, M1, part + s.off_w1 + br * H1 * F, nullptr, t, T);
  else s3_cross_tile_store(bwg, NT, ti, M1, N1, Kin1, part + s.off_w1, part + s.off_b1, t, T);
  DRGNN_PHASE3(15);
  } while (0);
  if (!valid && train && r == 0) {
#pragma unroll 1
    for (int i = t; i < s.n_params + 1; i += T) part[i] = 0.f;
  }
  // nobody leaves (or reuses its shared memory) while a peer may still read it
  if (CS > 1) cluster.sync();
  if (cta_times) g_cta_times[blockIdx.x][1] = s3_globaltimer();
  if (!P.fused_reduce || !train) return;
  s3_grid_reduce(s, C, scr, red);
  DRGNN_PHASE3(16);
}

// Reduction of the per-graph gradient rows (+ Adam) as its own launch: grids that are not co-resident.
static constexpr int RED3_SPLITS = 8;
__global__ void __launch_bounds__(32 * RED3_SPLITS) net_step_reduce_kernel(const drgnn_net_step_args s) {
  __shared__ float sh[3];
  __shared__ float psum[RED3_SPLITS][32];
  __shared__ bool is_last;
  const int lane = threadIdx.x & 31, q = threadIdx.x >> 5;
  const int e = blockIdx.x * 32 + lane;
  const int B = s.B, n = s.n_params;
  if (s.fuse_adam && threadIdx.x == 0) {
    const float st = s.step_dev[0] + 1.f;
    sh[0] = st;
    sh[1] = adam_bias_correction(s.beta1, st);
    sh[2] = adam_bias_correction(s.beta2, st);
  }
  {
    const int gs = (B + RED3_SPLITS - 1) / RED3_SPLITS;
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
    for (int u = 0; u < RED3_SPLITS; ++u) acc += psum[u][lane];
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
    __threadfence();
    unsigned* ticket = reinterpret_cast<unsigned*>(s.step_dev + 1);
    is_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    if (is_last) {
      *ticket = 0u;
      s.step_dev[0] = sh[0];
    }
  }
}
