// Loss (+ gradient w.r.t. the prediction), flat Adam, small utilities, library plumbing.
//
// Replaces the per-batch tail of NeuralNet._epoch (deeprank_gnn/NeuralNet.py:494-503):
// format_output (:616-631), MSELoss / weighted CrossEntropyLoss (:239-263) and
// torch.optim.Adam over 16 parameter tensors (:183-184, :503) by one launch each over flat
// buffers.  Nothing here reads back to the host, so the whole step can be replayed from a
// CUDA graph.
#include "common.cuh"

namespace drgnn {

const DeviceInfo& device_info() {
  static thread_local DeviceInfo info = {0, 0, -1};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
  if (info.device != dev) {
    int sms = 148, optin = 227 * 1024;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    info.sms = sms;
    info.smem_optin = optin;
    info.device = dev;
  }
  return info;
}

// single CTA: B_local is at most a few thousand
__global__ void __launch_bounds__(256) mse_loss_kernel(const float* __restrict__ pred, const float* __restrict__ y, int B,
                                                       float invB, int sigmoid, float* __restrict__ loss_out,
                                                       float* __restrict__ dpred) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    float p = pred[i], dp = 1.f;
    if (sigmoid) {
      p = 1.f / (1.f + expf(-p));
      dp = p * (1.f - p);
    }
    const float d = p - y[i];
    s += d * d;
    if (dpred) dpred[i] = 2.f * d * invB * dp;
  }
  s = warp_sum(s);
  if (lane_id() == 0) red[warp_id()] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    if (loss_out) loss_out[0] = t * invB;
  }
}

__global__ void __launch_bounds__(256) ce_loss_kernel(const float* __restrict__ logits, int ld,
                                                      const int64_t* __restrict__ target,
                                                      const float* __restrict__ class_w, int B, int nc, float inv_norm,
                                                      float* __restrict__ loss_out, float* __restrict__ dlogits) {
  __shared__ float red[8];
  float s = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float* z = logits + (int64_t)i * ld;
    float m = z[0];
    for (int c = 1; c < nc; ++c) m = fmaxf(m, z[c]);
    float se = 0.f;
    for (int c = 0; c < nc; ++c) se += expf(z[c] - m);
    const float lse = m + logf(se);
    const int t = (int)target[i];
    const float w = class_w ? class_w[t] : 1.f;
    s += w * (lse - z[t]);
    if (dlogits) {
      for (int c = 0; c < nc; ++c) {
        const float p = expf(z[c] - lse);
        dlogits[(int64_t)i * ld + c] = w * (p - (c == t ? 1.f : 0.f)) * inv_norm;
      }
    }
  }
  s = warp_sum(s);
  if (lane_id() == 0) red[warp_id()] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    if (loss_out) loss_out[0] = t * inv_norm;
  }
}

// torch.optim.Adam, single-tensor formulation (lerp, addcmul, sqrt / sqrt(bc2) + eps, addcdiv)
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                        float* __restrict__ m, float* __restrict__ v,
                                                        const float* __restrict__ step_dev, int64_t n, float lr, float beta1,
                                                        float beta2, float eps, float grad_scale) {
  __shared__ float sh[3];
  if (threadIdx.x == 0) {  // one thread does the double-precision powers (FP64 is slow on this part)
    const float st = step_dev[0] + 1.f;
    sh[0] = st;
    sh[1] = adam_bias_correction(beta1, st);
    sh[2] = adam_bias_correction(beta2, st);
  }
  __syncthreads();
  const float step = sh[0];
  const float step_size = lr / sh[1];
  const float bc2_sqrt = sqrtf(sh[2]);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * grad_scale;
    float mi = m[i], vi = v[i];
    mi = mi + (gi - mi) * (1.f - beta1);
    vi = vi * beta2 + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
}
__global__ void step_increment_kernel(float* step_dev) { step_dev[0] += 1.f; }

// Small parameter sets (every reference network: ~4-11 k floats): ONE block does the update and
// bumps the step counter itself (all threads read it, barrier, thread 0 writes) - one launch less.
__global__ void __launch_bounds__(1024) adam_flat_small_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                               float* __restrict__ m, float* __restrict__ v,
                                                               float* step_dev, int n, float lr, float beta1, float beta2,
                                                               float eps, float grad_scale) {
  __shared__ float sh[3];
  if (threadIdx.x == 0) {  // one thread does the double-precision powers (FP64 is slow on this part)
    const float st = step_dev[0] + 1.f;
    sh[0] = st;
    sh[1] = adam_bias_correction(beta1, st);
    sh[2] = adam_bias_correction(beta2, st);
  }
  __syncthreads();
  const float step = sh[0];
  const float step_size = lr / sh[1];
  const float bc2_sqrt = sqrtf(sh[2]);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float gi = g[i] * grad_scale;
    float mi = m[i], vi = v[i];
    mi = mi + (gi - mi) * (1.f - beta1);
    vi = vi * beta2 + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - step_size * (mi / denom);
  }
  __syncthreads();
  if (threadIdx.x == 0) step_dev[0] = step;
}

__global__ void __launch_bounds__(256) relu_mask_kernel(const float* __restrict__ g, int ldg, const float* __restrict__ out,
                                                        int ldo, int rows, const int32_t* rows_dev, int C,
                                                        float* __restrict__ gz, int ldgz) {
  const int n = rows_dev ? min(*rows_dev, rows) : rows;
  const int64_t total = (int64_t)n * C;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(idx / C), c = (int)(idx % C);
    gz[(int64_t)r * ldgz + c] = out[(int64_t)r * ldo + c] > 0.f ? g[(int64_t)r * ldg + c] : 0.f;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(T* p, T v, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace drgnn

using namespace drgnn;

extern "C" const char* drgnn_last_error(void) { return err_buf(); }
extern "C" int drgnn_version(void) { return 100; }
extern "C" int drgnn_device_sms(void) { return device_info().sms; }
extern "C" int drgnn_device_smem_optin(void) { return device_info().smem_optin; }

extern "C" int drgnn_mse_loss(const float* pred, const float* y, int32_t B_local, float inv_B_global, int32_t sigmoid,
                              float* loss_out, float* dpred, void* stream) {
  DRGNN_REQUIRE(B_local >= 0, "mse_loss: negative batch");
  DRGNN_REQUIRE(B_local == 0 || (pred && y), "mse_loss: NULL pointer");
  mse_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(pred, y, B_local, inv_B_global, sigmoid, loss_out, dpred);
  DRGNN_CHECK_LAUNCH("mse_loss_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_ce_loss(const float* logits, int32_t ld, const int64_t* target, const float* class_w, int32_t B_local,
                             int32_t n_classes, float inv_norm_global, float* loss_out, float* dlogits, void* stream) {
  DRGNN_REQUIRE(B_local >= 0 && n_classes > 0 && ld >= n_classes, "ce_loss: bad sizes");
  DRGNN_REQUIRE(B_local == 0 || (logits && target), "ce_loss: NULL pointer");
  ce_loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, ld, target, class_w, B_local, n_classes, inv_norm_global,
                                                      loss_out, dlogits);
  DRGNN_CHECK_LAUNCH("ce_loss_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* step_dev, int64_t n,
                               float lr, float beta1, float beta2, float eps, float grad_scale, void* stream) {
  DRGNN_REQUIRE(param && grad && exp_avg && exp_avg_sq && step_dev, "adam_flat: NULL pointer");
  DRGNN_REQUIRE(n >= 0, "adam_flat: negative size");
  cudaStream_t st = (cudaStream_t)stream;
  if (false && n > 0 && n <= 32768) {  // measured slower on B200 (12.5 us vs 4.2 + 1.9 us): one block serialises the memory round trips
    adam_flat_small_kernel<<<1, 1024, 0, st>>>(param, grad, exp_avg, exp_avg_sq, step_dev, (int)n, lr, beta1, beta2, eps,
                                               grad_scale);
    DRGNN_CHECK_LAUNCH("adam_flat_small_kernel");
    return DRGNN_OK;
  }
  if (n > 0) {
    const int blocks = (int)min_i64((n + 255) / 256, (int64_t)device_info().sms * 8);
    adam_flat_kernel<<<blocks, 256, 0, st>>>(param, grad, exp_avg, exp_avg_sq, step_dev, n, lr, beta1, beta2, eps, grad_scale);
    DRGNN_CHECK_LAUNCH("adam_flat_kernel");
  }
  step_increment_kernel<<<1, 1, 0, st>>>(step_dev);
  DRGNN_CHECK_LAUNCH("step_increment_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_relu_mask(const float* g, int32_t ldg, const float* out, int32_t ldo, int32_t rows,
                               const int32_t* rows_dev, int32_t C, float* gz, int32_t ldgz, void* stream) {
  DRGNN_REQUIRE(g && out && gz, "relu_mask: NULL pointer");
  DRGNN_REQUIRE(rows >= 0 && C > 0 && ldg >= C && ldo >= C && ldgz >= C, "relu_mask: bad sizes");
  if (rows == 0) return DRGNN_OK;
  const int64_t total = (int64_t)rows * C;
  const int blocks = (int)min_i64((total + 255) / 256, (int64_t)device_info().sms * 16);
  relu_mask_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(g, ldg, out, ldo, rows, rows_dev, C, gz, ldgz);
  DRGNN_CHECK_LAUNCH("relu_mask_kernel");
  return DRGNN_OK;
}

extern "C" int drgnn_fill_f32(float* p, float v, int64_t n, void* stream) {
  DRGNN_REQUIRE(p || n == 0, "fill_f32: NULL pointer");
  if (n <= 0) return DRGNN_OK;
  const int blocks = (int)min_i64((n + 255) / 256, (int64_t)device_info().sms * 16);
  fill_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(p, v, n);
  DRGNN_CHECK_LAUNCH("fill_kernel<float>");
  return DRGNN_OK;
}

extern "C" int drgnn_fill_i32(int32_t* p, int32_t v, int64_t n, void* stream) {
  DRGNN_REQUIRE(p || n == 0, "fill_i32: NULL pointer");
  if (n <= 0) return DRGNN_OK;
  const int blocks = (int)min_i64((n + 255) / 256, (int64_t)device_info().sms * 16);
  fill_kernel<int32_t><<<blocks, 256, 0, (cudaStream_t)stream>>>(p, v, n);
  DRGNN_CHECK_LAUNCH("fill_kernel<int32_t>");
  return DRGNN_OK;
}
