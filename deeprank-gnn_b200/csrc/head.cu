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

static constexpr int HD_ROWS = 32;    // rows per chunk
static constexpr int HD_THREADS = 512;

struct HeadSmem {
  int w1, w2, r, h, dh, dp, total;  // float offsets
  int ldw1, ldr, ldh;
};

__host__ __device__ inline HeadSmem head_plan(int C, int Hd, int out) {
  HeadSmem p;
  p.ldw1 = C + 1;   // W1s[h][k], stride C+1: conflict-free for thread-varying h and for thread-varying k
  p.ldr = C + 4;    // Rs[r][k]  (16-byte aligned rows)
  p.ldh = Hd + 4;   // Hs / dHs[r][h]
  int o = 0;
  p.w1 = o; o += Hd * p.ldw1;
  p.w2 = o; o += out * Hd;
  o = (o + 3) & ~3;
  p.r = o;  o += HD_ROWS * p.ldr;
  p.h = o;  o += HD_ROWS * p.ldh;
  p.dh = o; o += HD_ROWS * p.ldh;
  p.dp = o; o += HD_ROWS * (out + 1) * 2;   // pred and dpred of the chunk
  p.total = o;
  return p;
}

// acc[i] = sum_k A(m0+i, k) * Bv(k, n) for i < 8; one (8-row group, column) item per thread visit
template <typename FA, typename FB, typename FO>
__device__ __forceinline__ void mini_gemm(int M, int N, int K, FA a, FB b, FO out) {
  const int mgroups = (M + 7) >> 3;
  for (int item = threadIdx.x; item < mgroups * N; item += blockDim.x) {
    const int mg = item / N, n = item % N;
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int k = 0; k < K; ++k) {
      const float bv = b(k, n);
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(a(mg * 8 + i, k), bv, acc[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i)
      if (mg * 8 + i < M) out(mg * 8 + i, n, acc[i]);
  }
}

__global__ void __launch_bounds__(HD_THREADS, 1) head_kernel(const drgnn_head_args a) {
  extern __shared__ __align__(16) float hs[];
  __shared__ float red[HD_THREADS / 32];
  const int C = a.C, Hd = a.Hd, out = a.out, B = a.B;
  const HeadSmem P = head_plan(C, Hd, out);
  float* W1s = hs + P.w1;
  float* W2s = hs + P.w2;
  float* Rs = hs + P.r;
  float* Hs = hs + P.h;
  float* dHs = hs + P.dh;
  float* preds = hs + P.dp;
  float* dps = preds + HD_ROWS * (out + 1);
  const int t = threadIdx.x, T = blockDim.x;
  const bool bwd = a.task != 0 && a.dW1 != nullptr;
  const int ldw1 = P.ldw1, ldr = P.ldr, ldh = P.ldh, ldp = out + 1;

  for (int i = t; i < Hd * C; i += T) W1s[(i / C) * ldw1 + (i % C)] = a.W1[i];
  for (int i = t; i < out * Hd; i += T) W2s[i] = a.W2[i];
  float loss_acc = 0.f;

  for (int r0 = 0; r0 < B; r0 += HD_ROWS) {
    const int rows = min(HD_ROWS, B - r0);
    __syncthreads();  // weights staged / previous chunk fully consumed
    for (int i = t; i < HD_ROWS * C; i += T) {
      const int r = i / C, k = i % C;
      Rs[r * ldr + k] = r < rows ? a.R[(int64_t)(r0 + r) * a.ldr + k] : 0.f;
    }
    __syncthreads();
    // ---- H = relu(R W1^T + b1) * keep
    mini_gemm(HD_ROWS, Hd, C, [&](int m, int k) { return Rs[m * ldr + k]; },
              [&](int k, int n) { return W1s[n * ldw1 + k]; },
              [&](int m, int n, float v) {
                v += a.b1 ? __ldg(a.b1 + n) : 0.f;
                v = v < 0.f ? 0.f : v;
                if (a.keep && m < rows) v = a.keep[(int64_t)(r0 + m) * Hd + n] > 0.f ? v * a.keep_scale : 0.f;
                Hs[m * ldh + n] = m < rows ? v : 0.f;
                if (a.H && m < rows) a.H[(int64_t)(r0 + m) * Hd + n] = v;
              });
    __syncthreads();
    // ---- pred = H W2^T + b2
    mini_gemm(HD_ROWS, out, Hd, [&](int m, int k) { return Hs[m * ldh + k]; },
              [&](int k, int n) { return W2s[n * Hd + k]; },
              [&](int m, int n, float v) {
                v += a.b2 ? __ldg(a.b2 + n) : 0.f;
                preds[m * ldp + n] = v;
                if (m < rows) a.pred[(int64_t)(r0 + m) * out + n] = v;
              });
    __syncthreads();
    if (a.task == 0) continue;
    // ---- loss and dLoss/dpred of the chunk (one thread per row)
    if (t < HD_ROWS) {
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
    // ---- dW2 (+)= dpred^T H, db2 (+)= sum dpred      [out x Hd]
    mini_gemm(out, Hd, HD_ROWS, [&](int m, int k) { return dps[k * ldp + min(m, out - 1)]; },
              [&](int k, int n) { return Hs[k * ldh + n]; },
              [&](int m, int n, float v) {
                float* g = a.dW2 + (int64_t)m * Hd + n;
                *g = r0 == 0 ? v : *g + v;
              });
    if (t < out && a.db2) {
      float sacc = 0.f;
      for (int r = 0; r < HD_ROWS; ++r) sacc += dps[r * ldp + t];
      a.db2[t] = r0 == 0 ? sacc : a.db2[t] + sacc;
    }
    // ---- dH = (dpred W2) * (H > 0) * keep_scale
    mini_gemm(HD_ROWS, Hd, out, [&](int m, int k) { return dps[m * ldp + k]; },
              [&](int k, int n) { return W2s[k * Hd + n]; },
              [&](int m, int n, float v) { dHs[m * ldh + n] = Hs[m * ldh + n] > 0.f ? v * a.keep_scale : 0.f; });
    __syncthreads();
    // ---- dW1 (+)= dH^T R, db1 (+)= sum dH             [Hd x C]
    mini_gemm(Hd, C, HD_ROWS, [&](int m, int k) { return dHs[k * ldh + min(m, Hd - 1)]; },
              [&](int k, int n) { return Rs[k * ldr + n]; },
              [&](int m, int n, float v) {
                float* g = a.dW1 + (int64_t)m * C + n;
                *g = r0 == 0 ? v : *g + v;
              });
    if (a.db1) {
      for (int h = t; h < Hd; h += T) {
        float sacc = 0.f;
        for (int r = 0; r < HD_ROWS; ++r) sacc += dHs[r * ldh + h];
        a.db1[h] = r0 == 0 ? sacc : a.db1[h] + sacc;
      }
    }
    // ---- dR = dH W1                                   [rows x C]
    if (a.dR) {
      mini_gemm(HD_ROWS, C, Hd, [&](int m, int k) { return dHs[m * ldh + k]; },
                [&](int k, int n) { return W1s[k * ldw1 + n]; },
                [&](int m, int n, float v) {
                  if (m < rows) a.dR[(int64_t)(r0 + m) * a.lddr + n] = v;
                });
    }
  }
  if (a.task != 0 && a.loss) {
    float s = warp_sum(t < HD_ROWS ? loss_acc : 0.f);
    if (t == 0) a.loss[0] = s * a.inv_norm;   // only warp 0 holds row losses (HD_ROWS == 32)
  }
  (void)red;
}

}  // namespace drgnn

using namespace drgnn;

extern "C" int64_t drgnn_head_smem_bytes(int32_t C, int32_t Hd, int32_t out) {
  if (C <= 0 || Hd <= 0 || out <= 0) return DRGNN_ERR_INVALID;
  const int64_t bytes = (int64_t)head_plan(C, Hd, out).total * 4;
  if (bytes > device_info().smem_optin - 2048) return DRGNN_ERR_UNSUPPORTED;
  return bytes;
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
  if (smem < 0) return fail(DRGNN_ERR_UNSUPPORTED, "head: fc1 %d x %d does not fit shared memory", a->Hd, a->C);
  static thread_local int64_t configured = -1;
  if (smem > configured) {
    DRGNN_CHECK_CUDA(cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)device_info().smem_optin - 2048));
    configured = device_info().smem_optin - 2048;
  }
  head_kernel<<<1, HD_THREADS, smem, (cudaStream_t)stream>>>(*a);
  DRGNN_CHECK_LAUNCH("head_kernel");
  return DRGNN_OK;
}
