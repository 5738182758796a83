// Dense per-node transform on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
//   Y[r, 0:Fout] = act( X[r, 0:Fin] W + bias ) (* mask)         (include/drgnn.h section 3, math = 2)
//
// the nn.Linear / torch.mm of the reference layers (ginet.py:57-58, sGAT.py:73, foutnet.py:62-65) applied to
// node rows - the only GEMM-shaped work of the path (north_star: "tensor cores used only for the dense per-node
// feature x weight contraction").  fp32 parity is kept with the 3xTF32 split: x = x_hi + x_lo with both halves
// exactly representable in TF32, and D = A_hi B_hi + A_hi B_lo + A_lo B_hi accumulated in fp32 in TMEM
// (the dropped A_lo B_lo term is ~2^-22 relative).
//
// One CTA = 128 threads = one 128-row tile per iteration (persistent over tiles):
//   1. the tile's rows are split into A_hi / A_lo and written to shared memory in the canonical K-major
//      no-swizzle UMMA layout (core matrices of 8 rows x 16 bytes; LBO = stride between the two 16-byte
//      K chunks of an MMA, SBO = stride between 8-row groups); W is split once per CTA the same way;
//   2. ONE thread issues 3 x Fin/8 tcgen05.mma.kind::tf32 (M = 128, N = Fout, K = 8) into a TMEM
//      accumulator and commits them to an mbarrier;
//   3. every warp reads its 32 TMEM lanes (= 32 rows) with tcgen05.ld, applies bias / ReLU / mask and stores
//      whole output rows.
// Two CTAs per SM interleave (64-96 KB of shared memory, 64 of the 512 TMEM columns each), so one tile's
// loads and stores overlap the other's MMAs.  The kernel is HBM-bound by design (about 10 FLOP per byte at
// Fin = 32, Fout = 64); the tensor pipe only has to stay out of the way, which the fp32 FMA version cannot
// (75 TFLOP/s of FMA peak vs 6.5 TB/s x 10 FLOP/B).
#include "common.cuh"

namespace drgnn {

static constexpr int TC_THREADS = 128;
static constexpr int TC_BM = 128;
static constexpr int TC_TMEM_COLS = 64;
static constexpr int TC_MAX_KC = 16;          // Fin <= 64: at most 16 16-byte chunks per row

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t tc_f2tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
// x = hi + lo, hi = tf32(x), lo = tf32(x - hi)  (both with the 13 low mantissa bits zero)
__device__ __forceinline__ void tc_split(float x, float& hi, float& lo) {
  const uint32_t h = tc_f2tf32(x);
  hi = __uint_as_float(h);
  lo = __uint_as_float(tc_f2tf32(x - hi));
}
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE (cute::UMMA::SmemDescriptor): start address >> 4 in
// bits [0,14), leading byte offset >> 4 in [16,30), stride byte offset >> 4 in [32,46), version 1 in [46,48)
__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits [4,6) = 1), A = B = TF32 (bits [7,10) and
// [10,13) = 2), both K-major, N >> 3 in [17,23), M >> 4 in [24,29)
__host__ __device__ inline uint32_t tc_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "TC_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra TC_DONE_%=;\n"
      "bra TC_WAIT_%=;\n"
      "TC_DONE_%=:\n"
      "}\n" ::"r"(tc_smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// TMEM -> registers: this warp's 32 lanes x 16 / 32 consecutive columns (one row of the accumulator per thread)
__device__ __forceinline__ void tc_ld16(uint32_t (&v)[32], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t (&v)[32], uint32_t taddr) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}

__host__ __device__ inline int tc_a_region_words(int Fin, int Fout) {
  const int a = 2 * TC_BM * Fin, y = TC_BM * Fout;
  return a > y ? a : y;
}

// element (row, k) of a [rows x K] K-major operand tile in the canonical layout: core matrix (row / 8, k / 4)
// at ((k / 4) * (rows / 8) + row / 8) * 128 bytes, inside it row % 8 at 16-byte steps
__device__ __forceinline__ int tc_chunk_word(int row, int kchunk, int rows8) {
  return ((kchunk * rows8 + (row >> 3)) << 5) + ((row & 7) << 2);      // in 4-byte words
}

// diagnostic: cycles CTA 0 / thread 0 spent per phase, summed over its tiles (drgnn_debug_tc5_cycles):
// [0] split + store A, [1] fence + barrier + MMA issue, [2] prefetch issue, [3] wait for the MMAs, [7] TMEM read-back + staging, [4] output stores,
// [5] closing barrier, [6] tiles
__device__ unsigned long long g_tc5_phase[8];
#define TC5_T(i)                                                           \
  do {                                                                     \
    if (blockIdx.x == 0 && threadIdx.x == 0) {                             \
      const unsigned long long now_ = clock64();                           \
      g_tc5_phase[i] += now_ - tprev;                                      \
      tprev = now_;                                                        \
    }                                                                      \
  } while (0)

__global__ void __launch_bounds__(TC_THREADS, 3) linear_tcgen05_kernel(const drgnn_linear_args a, int n_tiles) {
  extern __shared__ __align__(128) float tsm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int t = threadIdx.x, warp = t >> 5;
  const int Fin = a.Fin, Fout = a.Fout;
  const int KC = Fin >> 2;                       // 16-byte K chunks per row
  const int rows = a.rows_dev ? min(*a.rows_dev, a.rows) : a.rows;
  float* Ahi = tsm;                              // [128 x Fin]; the region also stages the [128 x Fout] output tile
  float* Alo = Ahi + TC_BM * Fin;
  float* Bhi = tsm + tc_a_region_words(Fin, Fout);   // [Fout x Fin]
  float* Blo = Bhi + Fout * Fin;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_base_s)), "n"(TC_TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (t == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // W -> B_hi / B_lo as an [N = Fout][K = Fin] K-major operand (w_layout 0: W[o][k], 1: W[k][o])
  for (int idx = t; idx < Fout * KC; idx += TC_THREADS) {
    const int o = idx / KC, kc = idx - o * KC;
    float hi[4], lo[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = kc * 4 + j;
      const float w = a.w_layout == 0 ? __ldg(a.W + (int64_t)o * Fin + k) : __ldg(a.W + (int64_t)k * Fout + o);
      tc_split(w, hi[j], lo[j]);
    }
    const int wd = tc_chunk_word(o, kc, Fout >> 3);
    *reinterpret_cast<float4*>(Bhi + wd) = make_float4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<float4*>(Blo + wd) = make_float4(lo[0], lo[1], lo[2], lo[3]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = tmem_base_s;
  const uint32_t idesc = tc_idesc(TC_BM, Fout);
  const uint32_t a_lbo = (TC_BM >> 3) * 128, b_lbo = (uint32_t)(Fout >> 3) * 128;   // between the 16-byte K chunks
  uint32_t phase = 0;
  const uint64_t desc_ahi = tc_desc(tc_smem_u32(Ahi), a_lbo, 128), desc_alo = tc_desc(tc_smem_u32(Alo), a_lbo, 128);
  const uint64_t desc_bhi = tc_desc(tc_smem_u32(Bhi), b_lbo, 128), desc_blo = tc_desc(tc_smem_u32(Blo), b_lbo, 128);

  // software pipeline: the rows of the NEXT tile are fetched into registers before the epilogue of the current
  // one, so their global-memory latency hides behind the TMEM read-back and the output stores
  float4 xv[TC_MAX_KC];
  auto fetch = [&](int tile_) {
    const int r = tile_ * TC_BM + t;
    const float* xr = a.X + (int64_t)r * a.ldx;
#pragma unroll
    for (int kc = 0; kc < TC_MAX_KC; ++kc) {
      xv[kc] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kc < KC && tile_ < n_tiles && r < rows) xv[kc] = __ldg(reinterpret_cast<const float4*>(xr + kc * 4));
    }
  };
  fetch(blockIdx.x);
  unsigned long long tprev = clock64();
  if (blockIdx.x == 0 && t == 0) {
    for (int i = 0; i < 8; ++i) g_tc5_phase[i] = 0ull;
  }
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int r0 = tile * TC_BM;
    if (blockIdx.x == 0 && t == 0) { g_tc5_phase[6] += 1ull; tprev = clock64(); }
    // ---- 1. this thread's row -> A_hi / A_lo (rows past the end are zero)
#pragma unroll
    for (int kc = 0; kc < TC_MAX_KC; ++kc) {
      if (kc < KC) {
        const float4 v = xv[kc];
        float4 h, l;
        tc_split(v.x, h.x, l.x); tc_split(v.y, h.y, l.y); tc_split(v.z, h.z, l.z); tc_split(v.w, h.w, l.w);
        const int wd = tc_chunk_word(t, kc, TC_BM >> 3);
        *reinterpret_cast<float4*>(Ahi + wd) = h;
        *reinterpret_cast<float4*>(Alo + wd) = l;
      }
    }
    TC5_T(0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    // ---- 2. one thread issues the MMAs
    if (t == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      uint32_t acc = 0;
      // K = 8 per instruction = two 16-byte chunks: the descriptors of step ks are the base descriptors with the
      // start-address field (bits [0,14), units of 16 bytes) advanced by 2 ks LBO
      const uint64_t da = (uint64_t)((2 * a_lbo) >> 4), db = (uint64_t)((2 * b_lbo) >> 4);
      uint64_t d_alo = desc_alo, d_ahi = desc_ahi, d_blo = desc_blo, d_bhi = desc_bhi;
      for (int ks = 0; ks < (Fin >> 3); ++ks) {
        tc_mma(tmem_d, d_alo, d_bhi, idesc, acc);
        acc = 1;
        tc_mma(tmem_d, d_ahi, d_blo, idesc, acc);
        tc_mma(tmem_d, d_ahi, d_bhi, idesc, acc);
        d_alo += da; d_ahi += da; d_blo += db; d_bhi += db;
      }
      tc_commit(&bar);                                                // implies tcgen05.fence::before_thread_sync
    }
    TC5_T(1);
    fetch(tile + (int)gridDim.x);                                     // next tile's rows: in flight during the epilogue
    TC5_T(2);
    // ---- 3. epilogue: warp w owns TMEM lanes 32w .. 32w+31 = rows r0 + 32w + lane
    tc_mbar_wait(&bar, phase);
    phase ^= 1u;
    TC5_T(3);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // (a) TMEM -> registers -> shared memory.  The output tile [128 x Fout] is staged row-major in the (now idle)
    // operand region with the 16-byte chunks of a row XOR-swizzled by the row number: a warp's 32 rows hit all
    // banks evenly (4 wavefronts per 512-byte store, the minimum), and step (b) reads rows back conflict-free.
    float* Ys = Ahi;
    const int FC = Fout >> 2;                                         // 16-byte chunks per output row (4, 8 or 16)
    {
      // every TMEM load of the tile is issued before the single wait (their latencies overlap)
      uint32_t v0[32], v1[32];
      const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16);
      if (Fout >= 32) tc_ld32(v0, taddr); else tc_ld16(v0, taddr);
      if (Fout == 64) tc_ld32(v1, taddr + 32u);
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      const int n0 = Fout >= 32 ? 32 : 16;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        if (j < n0) {
          const int ch = (j >> 2) ^ (t & (FC - 1));
          *reinterpret_cast<float4*>(Ys + t * Fout + ch * 4) =
              make_float4(__uint_as_float(v0[j]), __uint_as_float(v0[j + 1]), __uint_as_float(v0[j + 2]), __uint_as_float(v0[j + 3]));
        }
      }
      if (Fout == 64) {
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const int ch = ((32 + j) >> 2) ^ (t & (FC - 1));
          *reinterpret_cast<float4*>(Ys + t * Fout + ch * 4) =
              make_float4(__uint_as_float(v1[j]), __uint_as_float(v1[j + 1]), __uint_as_float(v1[j + 2]), __uint_as_float(v1[j + 3]));
        }
      }
    }
    __syncthreads();
    TC5_T(7);
    // (b) shared memory -> global memory, coalesced: consecutive threads store consecutive 16-byte chunks of a row
    // (a warp writes whole 128-byte lines), bias / ReLU / mask applied on the way
    for (int c = t; c < TC_BM * FC; c += TC_THREADS) {
      const int rr = c / FC, ch = c - rr * FC;
      const int r = r0 + rr;
      if (r >= rows) continue;
      float4 x = *reinterpret_cast<const float4*>(Ys + rr * Fout + ((ch ^ (rr & (FC - 1))) << 2));
      const int o = ch << 2;
      if (a.bias) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + o));
        x.x += b.x; x.y += b.y; x.z += b.z; x.w += b.w;
      }
      if (a.relu) {                                                   // v < 0 ? 0 : v keeps NaN like torch.relu
        x.x = x.x < 0.f ? 0.f : x.x; x.y = x.y < 0.f ? 0.f : x.y; x.z = x.z < 0.f ? 0.f : x.z; x.w = x.w < 0.f ? 0.f : x.w;
      }
      if (a.out_mask) {
        const float* mk = a.out_mask + (int64_t)r * a.ld_mask + o;
        x.x = mk[0] > 0.f ? x.x * a.mask_scale : 0.f; x.y = mk[1] > 0.f ? x.y * a.mask_scale : 0.f;
        x.z = mk[2] > 0.f ? x.z * a.mask_scale : 0.f; x.w = mk[3] > 0.f ? x.w * a.mask_scale : 0.f;
      }
      *reinterpret_cast<float4*>(a.Y + (int64_t)r * a.ldy + o) = x;
    }
    TC5_T(4);
    // the next tile overwrites the operand tiles and the accumulator: everybody is done reading them
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    TC5_T(5);
  }
  if (warp == 0)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "n"(TC_TMEM_COLS) : "memory");
}

}  // namespace drgnn

using namespace drgnn;

// 0 when the tcgen05 kernel takes this shape (else the caller uses the mma.sync / FMA kernels)
extern "C" int drgnn_linear_tcgen05_supported(const drgnn_linear_args* a) {
  if (a == nullptr) return 0;
  const bool ok = a->groups == 1 && a->Fin % 8 == 0 && a->Fin >= 8 && a->Fin <= 64 &&
                  (a->Fout == 16 || a->Fout == 32 || a->Fout == 64) &&      // N % 16, power of two (swizzled output staging)
                  a->Fout <= TC_TMEM_COLS && a->ldx % 4 == 0 && a->ldy % 4 == 0 && ((uintptr_t)a->X % 16) == 0 &&
                  ((uintptr_t)a->Y % 16) == 0 && (a->bias == nullptr || ((uintptr_t)a->bias % 16) == 0);
  return ok ? 1 : 0;
}

extern "C" int drgnn_linear_tcgen05(const drgnn_linear_args* a, void* stream) {
  DRGNN_REQUIRE(a != nullptr && a->X && a->W && a->Y, "linear_tcgen05: NULL pointer");
  DRGNN_REQUIRE(drgnn_linear_tcgen05_supported(a), "linear_tcgen05: unsupported shape (groups 1, Fin %% 8, Fin <= 64, Fout %% 16, Fout <= 64, "
                                                   "16-byte aligned rows)");
  if (a->rows == 0) return DRGNN_OK;
  const int n_tiles = (a->rows + TC_BM - 1) / TC_BM;
  const size_t smem = (size_t)4 * (tc_a_region_words(a->Fin, a->Fout) + 2 * a->Fout * a->Fin);
  static thread_local size_t configured = 0;
  if (smem > configured) {
    DRGNN_CHECK_CUDA(cudaFuncSetAttribute(linear_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(96 * 1024)));
    configured = 96 * 1024;
  }
  int per_sm = (int)((device_info().smem_optin + 1024) / (smem + 1024));     // CTAs one SM holds (shared memory; TMEM: 8)
  if (per_sm > 3) per_sm = 3;
  if (per_sm < 1) per_sm = 1;
  int grid = per_sm * device_info().sms;
  if (grid > n_tiles) grid = n_tiles;
  linear_tcgen05_kernel<<<grid, TC_THREADS, smem, (cudaStream_t)stream>>>(*a, n_tiles);
  DRGNN_CHECK_LAUNCH("linear_tcgen05_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_debug_tc5_cycles(uint64_t* out8) {
  DRGNN_REQUIRE(out8 != nullptr, "debug_tc5_cycles: NULL");
  DRGNN_CHECK_CUDA(cudaMemcpyFromSymbol(out8, g_tc5_phase, sizeof(unsigned long long) * 8));
  return DRGNN_OK;
}
