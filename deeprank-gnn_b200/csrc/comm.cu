// Multi-GPU gradient exchange fused with the optimiser: the path's only collective.
//
// The reference is single-device; data-parallel training needs, per step, sum_over_ranks(flat
// gradient | loss) followed by Adam (NeuralNet.py:502-503 on every rank).  Instead of
// [reduce kernel] -> ncclAllReduce -> [Adam kernel] this file does all three in ONE launch over
// NVLink peer memory:
//
//   every block: local value of its 256 elements (sum of the per-graph partial rows in graph order,
//                or the already reduced local gradient)
//             -> plain stores of the chunk into EVERY rank's exchange buffer (slot = my rank)
//             -> st.release.sys of the block's flag in every rank's flag array
//             -> spin (ld.acquire.sys) until the same block of every rank has delivered
//             -> sum the `world` slots in RANK ORDER (identical on all ranks => weights stay
//                bit-identical) -> Adam on the element.
//
// Blocks only depend on the SAME block index of the peers, so there is no grid-wide barrier and no
// co-residency requirement.  Exchange buffers and flags are double-buffered on the parity of a
// device-side epoch counter: a rank can be at most one step ahead of a peer that is still reading
// (its step e+1 needs that peer's step-e+1 flags), so two parities suffice.  A watchdog turns a
// lost peer into an error status (ctr[2]) instead of a hung GPU.
#include <string.h>

#include "common.cuh"

namespace drgnn {

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

static constexpr int PR_THREADS = 256;

__global__ void __launch_bounds__(PR_THREADS) peer_reduce_adam_kernel(const drgnn_peer_comm c, const drgnn_peer_adam_args a) {
  __shared__ float sh[3];
  __shared__ uint32_t s_epoch;
  const int t = threadIdx.x;
  const int e = blockIdx.x * PR_THREADS + t;
  if (t == 0) {
    s_epoch = *reinterpret_cast<volatile uint32_t*>(c.ctr) + 1u;
    const float st = a.step_dev ? a.step_dev[0] + 1.f : 1.f;
    sh[0] = st;
    sh[1] = adam_bias_correction(a.beta1, st);
    sh[2] = adam_bias_correction(a.beta2, st);
  }
  __syncthreads();
  const uint32_t epoch = s_epoch;
  const int par = (int)(epoch & 1u);
  const int world = c.world, rank = c.rank;
  // ---- local value
  float acc = 0.f;
  if (e < a.n_sum) {
    if (a.partial) {
      if (e <= a.n_params) {     // rows carry [gradients | loss term]; the padding behind them is not summed
#pragma unroll 8
        for (int g = 0; g < a.B; ++g) acc += a.partial[(int64_t)g * a.partial_ld + e];
      }
    } else {
      acc = a.grads[e];
    }
    // ---- deliver to every rank (own slot included: one code path)
    const int64_t slot = ((int64_t)par * world + rank) * c.stride + e;
    for (int p = 0; p < world; ++p) c.xbuf[(rank + p) % world][slot] = acc;
  }
  __syncthreads();
  if (t < world) {
    __threadfence_system();
    st_release_sys(c.xflag[t] + ((int64_t)par * world + rank) * c.max_blocks + blockIdx.x, epoch);
    // ---- wait for block `blockIdx.x` of rank t
    const uint32_t* f = c.xflag[rank] + ((int64_t)par * world + t) * c.max_blocks + blockIdx.x;
    const uint64_t t0 = global_ns(), limit = c.timeout_ns ? c.timeout_ns : 20000000000ull;
    while (ld_acquire_sys(f) != epoch) {
      if (global_ns() - t0 > limit) {
        atomicOr(c.ctr + 2, 1u);
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
  // ---- rank-ordered sum + Adam
  if (e < a.n_sum) {
    float tot = 0.f;
    const float* mine = c.xbuf[rank] + (int64_t)par * world * c.stride + e;
    for (int r = 0; r < world; ++r) tot += __ldcg(mine + (int64_t)r * c.stride);
    a.grads[e] = tot;
    if (a.apply_adam && e < a.n_params) {
      float mi = a.adam_m[e], vi = a.adam_v[e];
      mi = mi + (tot - mi) * (1.f - a.beta1);
      vi = vi * a.beta2 + (1.f - a.beta2) * tot * tot;
      a.adam_m[e] = mi;
      a.adam_v[e] = vi;
      const float denom = sqrtf(vi) / sqrtf(sh[2]) + a.eps;
      a.adam_p[e] = a.adam_p[e] - (a.lr / sh[1]) * (mi / denom);
    }
  }
  __syncthreads();
  if (t == 0) {
    // the last block to finish publishes the new epoch / Adam step: no block of this launch can see it
    if (atomicAdd(c.ctr + 1, 1u) == gridDim.x - 1) {
      c.ctr[1] = 0u;
      *reinterpret_cast<volatile uint32_t*>(c.ctr) = epoch;
      if (a.apply_adam && a.step_dev) a.step_dev[0] = sh[0];
    }
  }
}

}  // namespace drgnn

using namespace drgnn;

extern "C" int drgnn_comm_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64) {
  DRGNN_REQUIRE(bytes > 0 && dev_ptr && handle64, "comm_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == DRGNN_IPC_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  DRGNN_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
  DRGNN_CHECK_CUDA(cudaMemset(p, 0, (size_t)bytes));
  DRGNN_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return fail(DRGNN_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
  }
  memcpy(handle64, &h, sizeof(h));
  *dev_ptr = p;
  return DRGNN_OK;
}

extern "C" int drgnn_comm_open(const unsigned char* handle64, void** peer_ptr) {
  DRGNN_REQUIRE(handle64 && peer_ptr, "comm_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  DRGNN_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *peer_ptr = p;
  return DRGNN_OK;
}

extern "C" int drgnn_comm_close(void* peer_ptr) {
  if (peer_ptr) DRGNN_CHECK_CUDA(cudaIpcCloseMemHandle(peer_ptr));
  return DRGNN_OK;
}

extern "C" int drgnn_comm_free(void* dev_ptr) {
  if (dev_ptr) DRGNN_CHECK_CUDA(cudaFree(dev_ptr));
  return DRGNN_OK;
}

extern "C" int drgnn_comm_status(const void* region, uint32_t* ctr4) {
  DRGNN_REQUIRE(region && ctr4, "comm_status: bad arguments");
  DRGNN_CHECK_CUDA(cudaMemcpy(ctr4, region, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost));   // synchronises
  return DRGNN_OK;
}

extern "C" int drgnn_peer_reduce_adam(const drgnn_peer_comm* c, const drgnn_peer_adam_args* a, void* stream) {
  DRGNN_REQUIRE(c && a, "peer_reduce_adam: NULL arguments");
  DRGNN_REQUIRE(c->world >= 1 && c->world <= DRGNN_MAX_PEERS && c->rank >= 0 && c->rank < c->world,
                "peer_reduce_adam: bad world / rank %d / %d", c->world, c->rank);
  DRGNN_REQUIRE(c->ctr && c->stride >= a->n_sum && c->max_blocks > 0, "peer_reduce_adam: bad exchange layout");
  for (int r = 0; r < c->world; ++r)
    DRGNN_REQUIRE(c->xbuf[r] && c->xflag[r], "peer_reduce_adam: rank %d has no exchange buffer", r);
  DRGNN_REQUIRE(a->grads && a->n_sum > 0 && a->n_params >= 0 && a->n_params <= a->n_sum, "peer_reduce_adam: bad sizes");
  DRGNN_REQUIRE(!a->partial || (a->B >= 0 && a->partial_ld >= a->n_sum), "peer_reduce_adam: bad partial rows");
  DRGNN_REQUIRE(!a->apply_adam || (a->adam_p && a->adam_m && a->adam_v && a->step_dev), "peer_reduce_adam: Adam buffers missing");
  const int blocks = (a->n_sum + PR_THREADS - 1) / PR_THREADS;
  DRGNN_REQUIRE(blocks <= c->max_blocks, "peer_reduce_adam: %d blocks exceed the flag capacity %d", blocks, c->max_blocks);
  peer_reduce_adam_kernel<<<blocks, PR_THREADS, 0, (cudaStream_t)stream>>>(*c, *a);
  DRGNN_CHECK_LAUNCH("peer_reduce_adam_kernel");
  return DRGNN_OK;
}
