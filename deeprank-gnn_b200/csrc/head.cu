// Fused network head: fc1 (+ReLU, +dropout) -> fc2 -> loss -> backward of both layers, ONE launch.
//
// Replaces, for the B graph read-outs of a mini-batch, the tail of the reference forward
// (deeprank_gnn/ginet.py:136-139, sGAT.py:134-135, foutnet.py:121-122), format_output + loss
// (NeuralNet.py:616-631, 239-263, 500) and the autograd backward of those ops: nine small
// launches in the op-by-op path (two transforms, loss, two weight gradients with their
// reductions, two input gradients).  All of it works on B rows (B = graphs per batch, 64 in the
// headline configuration), so it is pure launch latency; one CTA does it out of shared memory.
//
//   H    = relu(R W1^T + b1) (* keep * keep_scale)         [B, Hd]
//   pred = H W2^T + b2                                      [B, out]
//   loss, dpred  (MSE, MSE of sigmoid, or class-weighted cross entropy; task 0 = forward only)
//   dW2 = dpred^T H, db2 = sum dpred, dH = (dpred W2) * (H > 0) * keep_scale
//   dW1 = dH^T R,    db1 = sum dH,    dR = dH W1
// Rows are processed in chunks of 32; weight gradients accumulate in global memory in chunk
// order (single CTA => deterministic).
#include "common.cuh"

namespace drgnn {

static constexpr int HD_THREADS = 512;

// shared-memory plan (float offsets).  Operands are kept in the layout each product wants:
//   C[m][n] = sum_k At[k][m] * Bm[k][n]   with an 8 (m) x 4 (n) register tile per thread, so every
// k step is two 16-byte loads of At (broadcast inside a warp), one of Bm and 32 FMAs.
struct HeadSmem {
  int w1, w1t, w2, r, rt, h, dh, dht, dp, total;
  int RB;  // rows per chunk (64, or 32 when shared memory is short)
};

__host__ __device__ inline HeadSmem head_plan(int C, int Hd, int out, int RB) {
  HeadSmem p;
  p.RB = RB;
  int o = 0;
  p.w1 = o;  o += Hd * C;        // W1 [Hd][C]    (Bm of dR = dH W1)
  p.w1t = o; o += C * (Hd + 4);  // W1^T [C][Hd+4] (Bm of H = R W1^T; +4: transposing stores spread over banks)
  p.w2 = o;  o += ((out * Hd + 3) & ~3);
  p.r = o;   o += RB * C;        // R [r][c]      (Bm of dW1 = dH^T R)
  p.rt = o;  o += C * (RB + 4);  // R^T [c][RB+4] (At of H)
  p.h = o;   o += RB * Hd;       // H [r][h]      (ReLU / dropout mask, fc2, dW2)
  p.dh = o;  o += RB * Hd;       // dH [r][h]     (At of dW1)
  p.dht = o; o += Hd * (RB + 4); // dH^T [h][RB+4] (At of dR)
  p.dp = o;  o += 2 * RB * (out + 1);
  p.total = o;
  return p;
}

__global__ void __launch_bounds__(HD_THREADS, 1) head_kernel(const drgnn_head_args a, int RB) {
  extern __shared__ __align__(16) float hs[];
  const int C = a.C, Hd = a.Hd, out = a.out, B = a.B;
  const HeadSmem P = head_plan(C, Hd, out, RB);
  float* W1s = hs + P.w1;
  float* W1t = hs + P.w1t;
  float* W2s = hs + P.w2;
  float* Rs = hs + P.r;
  float* Rt = hs + P.rt;
  float* Hs = hs + P.h;
  float* dHs = hs + P.dh;
  float* dHt = hs + P.dht;
  float* preds = hs + P.dp;
  const int ldp = out + 1;
  float* dps = preds + RB * ldp;
  const int t = threadIdx.x, T = blockDim.x;
  const int lane = t & 31, warp = t >> 5, nwarps = T >> 5;
  const bool bwd = a.task != 0 && a.dW1 != nullptr;

  for (int i = t; i < Hd * C; i += T) {
    const float w = a.W1[i];
    const int h = i / C, c = i - h * C;
    W1s[i] = w;
    W1t[c * (Hd + 4) + h] = w;
  }
  for (int i = t; i < out * Hd; i += T) W2s[i] = a.W2[i];
  float loss_acc = 0.f;  // thread r < RB accumulates the loss of row r of every chunk

  for (int r0 = 0; r0 < B; r0 += RB) {
    const int rows = min(RB, B - r0);
    __syncthreads();  // weights staged / previous chunk fully consumed
    for (int i = t; i < RB * C; i += T) {
      const int r = i / C, c = i - r * C;
      const float v = r < rows ? a.R[(int64_t)(r0 + r) * a.ldr + c] : 0.f;
      Rs[i] = v;
      Rt[c * (RB + 4) + r] = v;
    }
    __syncthreads();
    // ---- H = relu(R W1^T + b1) * keep                                   [RB x Hd]
    tile_gemm(Rt, RB + 4, W1t, Hd + 4, RB, Hd, C, [&](int m, int n, float v) {
      v += a.b1 ? __ldg(a.b1 + n) : 0.f;
      v = v < 0.f ? 0.f : v;
      if (m < rows) {
        if (a.keep) v = a.keep[(int64_t)(r0 + m) * Hd + n] > 0.f ? v * a.keep_scale : 0.f;
        if (a.H) a.H[(int64_t)(r0 + m) * Hd + n] = v;
      } else {
        v = 0.f;
      }
      Hs[m * Hd + n] = v;
    });
    __syncthreads();
    // ---- pred = H W2^T + b2: one warp per (row, output), lanes over the hidden units
    for (int item = warp; item < RB * out; item += nwarps) {
      const int r = item / out, o = item - r * out;
      float sacc = 0.f;
      for (int h = lane; h < Hd; h += 32) sacc = fmaf(Hs[r * Hd + h], W2s[o * Hd + h], sacc);
      sacc = warp_sum(sacc);
      if (lane == 0) {
        sacc += a.b2 ? __ldg(a.b2 + o) : 0.f;
        preds[r * ldp + o] = sacc;
        if (r < rows) a.pred[(int64_t)(r0 + r) * out + o] = sacc;
      }
    }
    __syncthreads();
    if (a.task == 0) continue;
    // ---- loss and dLoss/dpred of the chunk (one thread per row)
    if (t < RB) {
      const int r = t;
      if (r < rows) {
        if (a.task == 3) {
          const float* z = preds + r * ldp;
          float mx = z[0];
          for (int c = 1; c < out; ++c) mx = fmaxf(mx, z[c]);
          float se = 0.f;
          for (int c = 0; c < out; ++c) se += expf(z[c] - mx);
          const float lse = mx + logf(se);
          const int tc = (int)a.y_class[r0 + r];
          const float w = a.class_w ? a.class_w[tc] : 1.f;
          loss_acc += w * (lse - z[tc]);
          for (int c = 0; c < out; ++c)
            dps[r * ldp + c] = w * (expf(z[c] - lse) - (c == tc ? 1.f : 0.f)) * a.inv_norm;
        } else {
          for (int c = 0; c < out; ++c) {
            float p = preds[r * ldp + c], dp = 1.f;
            if (a.task == 2) {
              p = 1.f / (1.f + expf(-p));
              dp = p * (1.f - p);
            }
            const float d = p - a.y[(int64_t)(r0 + r) * out + c];
            loss_acc += d * d;
            dps[r * ldp + c] = 2.f * d * a.inv_norm * dp;
          }
        }
      } else {
        for (int c = 0; c < out; ++c) dps[r * ldp + c] = 0.f;
      }
    }
    __syncthreads();
    if (!bwd) continue;
    // ---- dW2 (+)= dpred^T H, db2 (+)= sum dpred: one thread per (output, hidden unit)
    for (int item = t; item < out * Hd; item += T) {
      const int o = item / Hd, h = item - o * Hd;
      float sacc = 0.f;
      for (int r = 0; r < RB; ++r) sacc = fmaf(dps[r * ldp + o], Hs[r * Hd + h], sacc);
      float* g = a.dW2 + item;
      *g = r0 == 0 ? sacc : *g + sacc;
    }
    if (t < out && a.db2) {
      float sacc = 0.f;
      for (int r = 0; r < RB; ++r) sacc += dps[r * ldp + t];
      a.db2[t] = r0 == 0 ? sacc : a.db2[t] + sacc;
    }
    // ---- dH = (dpred W2) * (H > 0) * keep_scale, stored in both layouts
    for (int item = t; item < RB * Hd; item += T) {
      const int r = item / Hd, h = item - r * Hd;
      float sacc = 0.f;
      for (int o = 0; o < out; ++o) sacc = fmaf(dps[r * ldp + o], W2s[o * Hd + h], sacc);
      sacc = Hs[item] > 0.f ? sacc * a.keep_scale : 0.f;
      dHs[item] = sacc;
      dHt[h * (RB + 4) + r] = sacc;
    }
    __syncthreads();
    // ---- dW1 (+)= dH^T R                                                [Hd x C]
    tile_gemm(dHs, Hd, Rs, C, Hd, C, RB, [&](int m, int n, float v) {
      float* g = a.dW1 + (int64_t)m * C + n;
      *g = r0 == 0 ? v : *g + v;
    });
    if (a.db1) {
      for (int h = t; h < Hd; h += T) {
        float sacc = 0.f;
        for (int r = 0; r < RB; ++r) sacc += dHs[r * Hd + h];
        a.db1[h] = r0 == 0 ? sacc : a.db1[h] + sacc;
      }
    }
    // ---- dR = dH W1                                                     [RB x C]
    if (a.dR) {
      tile_gemm(dHt, RB + 4, W1s, C, RB, C, Hd, [&](int m, int n, float v) {
        if (m < rows) a.dR[(int64_t)(r0 + m) * a.lddr + n] = v;
      });
    }
  }
  if (a.task != 0 && a.loss) {
    // rows live in threads 0..RB-1 (warps 0 and 1 when RB = 64)
    __shared__ float lred[2];
    const float s = warp_sum(t < RB ? loss_acc : 0.f);
    if (lane == 0 && warp < 2) lred[warp] = s;
    __syncthreads();
    if (t == 0) a.loss[0] = (lred[0] + lred[1]) * a.inv_norm;
  }
}

static inline int head_rows_per_chunk(int C, int Hd, int out) {
  const int64_t budget = (int64_t)device_info().smem_optin - 2048;
  if ((int64_t)head_plan(C, Hd, out, 64).total * 4 <= budget) return 64;
  if ((int64_t)head_plan(C, Hd, out, 32).total * 4 <= budget) return 32;
  return 0;
}

}  // namespace drgnn

using namespace drgnn;

extern "C" int64_t drgnn_head_smem_bytes(int32_t C, int32_t Hd, int32_t out) {
  if (C <= 0 || Hd <= 0 || out <= 0) return DRGNN_ERR_INVALID;
  if (C % 4 != 0 || Hd % 8 != 0) return DRGNN_ERR_UNSUPPORTED;
  const int rb = head_rows_per_chunk(C, Hd, out);
  if (rb == 0) return DRGNN_ERR_UNSUPPORTED;
  return (int64_t)head_plan(C, Hd, out, rb).total * 4;
}

extern "C" int drgnn_head(const drgnn_head_args* a, void* stream) {
  DRGNN_REQUIRE(a != nullptr, "head: args is NULL");
  DRGNN_REQUIRE(a->B >= 0 && a->C > 0 && a->Hd > 0 && a->out > 0, "head: bad sizes");
  DRGNN_REQUIRE(a->R && a->W1 && a->W2 && a->pred, "head: NULL pointer");
  DRGNN_REQUIRE(a->task >= 0 && a->task <= 3, "head: bad task %d", a->task);
  DRGNN_REQUIRE(a->task == 0 || a->task == 3 || a->y, "head: regression needs y");
  DRGNN_REQUIRE(a->task != 3 || a->y_class, "head: classification needs y_class");
  DRGNN_REQUIRE(a->dW1 == nullptr || (a->dW2 && a->task != 0), "head: backward needs dW1, dW2 and a loss");
  DRGNN_REQUIRE(a->ldr >= a->C && (!a->dR || a->lddr >= a->C), "head: leading dimension too small");
  if (a->B == 0) return DRGNN_OK;
  const int64_t smem = drgnn_head_smem_bytes(a->C, a->Hd, a->out);
  if (smem < 0)
    return fail(DRGNN_ERR_UNSUPPORTED, "head: fc1 %d x %d does not fit one CTA (needs C %% 4 == 0, Hd %% 8 == 0)", a->Hd, a->C);
  const int rb = head_rows_per_chunk(a->C, a->Hd, a->out);
  static thread_local int64_t configured = -1;
  if (smem > configured) {
    DRGNN_CHECK_CUDA(cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)device_info().smem_optin - 2048));
    configured = device_info().smem_optin - 2048;
  }
  head_kernel<<<1, HD_THREADS, smem, (cudaStream_t)stream>>>(*a, rb);
  DRGNN_CHECK_LAUNCH("head_kernel");
  return DRGNN_OK;
}
