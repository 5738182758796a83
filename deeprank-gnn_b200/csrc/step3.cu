// C-ABI of the general cluster whole-step kernel (include/drgnn.h section 8b, kernel in fused_step3.cuh):
// GINet / sGAT / FoutNet, one graph per thread-block cluster, the node dimension tiled over the cluster's CTAs.
#include <cooperative_groups.h>
#include <float.h>
#include <limits.h>
#include <string.h>

#include "common.cuh"

namespace drgnn {
#include "tc_tiles.cuh"
#include "fused_step3.cuh"
}  // namespace drgnn

using namespace drgnn;

static bool step3_dims_ok(int kind, int F, int h1, int h2, int Hd, int out) {
  return kind >= 0 && kind <= 2 && F > 0 && h1 > 0 && h2 > 0 && Hd > 0 && out > 0 && F % 4 == 0 && h1 % 4 == 0 && h2 % 4 == 0 &&
         ((kind == 0 ? 2 : 1) * h2) % 4 == 0;
}

extern "C" int64_t drgnn_net_step_smem_bytes_l(int32_t kind, int32_t tiles, int32_t F, int32_t h1, int32_t h2, int32_t max_n,
                                               int32_t max_k, int32_t max_q, int32_t max_e, int32_t Hd, int32_t out, int32_t layers3) {
  if (max_n <= 0 || max_k <= 0 || max_q <= 0 || max_e <= 0) return DRGNN_ERR_INVALID;
  if (!step3_dims_ok(kind, F, h1, h2, Hd, out)) return DRGNN_ERR_UNSUPPORTED;
  if (tiles != 1 && tiles != 2 && tiles != 4 && tiles != 8) return DRGNN_ERR_INVALID;
  if (tiles * (kind == 0 ? 2 : 1) > 8) return DRGNN_ERR_UNSUPPORTED;          // portable cluster size
  if ((int64_t)max_n * (2 * F + h1 + 8) > (1 << 24) || max_e > (1 << 24) || (int64_t)Hd * h2 > (1 << 22)) return DRGNN_ERR_UNSUPPORTED;
  // staged (the graph's blob + feature tile in every CTA's shared memory) when that fits, else streamed from L2
  if (layers3 && kind == 0) return DRGNN_ERR_UNSUPPORTED;
  for (int stage = 1; stage >= 0; --stage) {
    const int64_t bytes = 4 * (int64_t)step3_plan(kind, tiles, stage, F, h1, h2, max_n, max_k, max_q, max_e, Hd, out, layers3).total;
    if (bytes <= device_info().smem_optin - 1024) return bytes;
  }
  return DRGNN_ERR_UNSUPPORTED;
}
extern "C" int64_t drgnn_net_step_smem_bytes(int32_t kind, int32_t tiles, int32_t F, int32_t h1, int32_t h2, int32_t max_n,
                                             int32_t max_k, int32_t max_q, int32_t max_e, int32_t Hd, int32_t out) {
  return drgnn_net_step_smem_bytes_l(kind, tiles, F, h1, h2, max_n, max_k, max_q, max_e, Hd, out, 0);
}

static int step3_stage(int kind, int tiles, int F, int h1, int h2, int max_n, int max_k, int max_q, int max_e, int Hd, int out,
                       int layers3) {
  const int64_t bytes = 4 * (int64_t)step3_plan(kind, tiles, 1, F, h1, h2, max_n, max_k, max_q, max_e, Hd, out, layers3).total;
  return bytes <= device_info().smem_optin - 1024 ? 1 : 0;
}

extern "C" int drgnn_net_step_pick_tiles_l(int32_t kind, int32_t F, int32_t h1, int32_t h2, int32_t max_n, int32_t max_k,
                                           int32_t max_q, int32_t max_e, int32_t Hd, int32_t out, int32_t layers3) {
  for (int tiles = 1; tiles <= 8; tiles *= 2) {
    const int64_t b = drgnn_net_step_smem_bytes_l(kind, tiles, F, h1, h2, max_n, max_k, max_q, max_e, Hd, out, layers3);
    if (b >= 0) return tiles;
    if (b == DRGNN_ERR_INVALID) return DRGNN_ERR_INVALID;
  }
  return DRGNN_ERR_UNSUPPORTED;
}
extern "C" int drgnn_net_step_pick_tiles(int32_t kind, int32_t F, int32_t h1, int32_t h2, int32_t max_n, int32_t max_k,
                                         int32_t max_q, int32_t max_e, int32_t Hd, int32_t out) {
  return drgnn_net_step_pick_tiles_l(kind, F, h1, h2, max_n, max_k, max_q, max_e, Hd, out, 0);
}

static int step3_configure(int64_t smem) {
  static thread_local int64_t configured = -1;
  if (smem > configured) {
    DRGNN_CHECK_CUDA(cudaFuncSetAttribute(net_graph_step3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)device_info().smem_optin - 1024));
    configured = device_info().smem_optin - 1024;
  }
  return DRGNN_OK;
}

static int step3_max_clusters(int cs, int64_t smem) {
  static thread_local int64_t c_smem = -1;
  static thread_local int c_cs = -1, cached = 0;
  if (c_smem == smem && c_cs == cs) return cached;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(cs * device_info().sms));
  cfg.blockDim = dim3(S3_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int nc = 0;
  const cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, net_graph_step3_kernel, &cfg);
  if (e != cudaSuccess) {
    fail(DRGNN_ERR_CUDA, "cudaOccupancyMaxActiveClusters(net_graph_step3_kernel, cluster %d): %s", cs, cudaGetErrorString(e));
    (void)cudaGetLastError();
    nc = 0;
  }
  cached = nc;
  c_smem = smem;
  c_cs = cs;
  return nc;
}

extern "C" int drgnn_net_step_max_clusters(int32_t kind, int32_t tiles, int64_t smem_bytes) {
  if (kind < 0 || kind > 2 || tiles < 1 || tiles > 8 || smem_bytes < 0 || smem_bytes > device_info().smem_optin - 1024)
    return DRGNN_ERR_INVALID;
  if (step3_configure(smem_bytes) != DRGNN_OK) return DRGNN_ERR_CUDA;
  return step3_max_clusters(tiles * (kind == 0 ? 2 : 1), smem_bytes);
}

static thread_local int g_step3_launches = 0;
static thread_local int g_step3_tiles = 0;
extern "C" int drgnn_net_step_last_launches(void) { return g_step3_launches; }
extern "C" int drgnn_net_step_last_tiles(void) { return g_step3_tiles; }

extern "C" int drgnn_net_step(const drgnn_net_step_args* s, void* stream) {
  DRGNN_REQUIRE(s != nullptr, "net_step: args is NULL");
  DRGNN_REQUIRE(s->B >= 0, "net_step: negative batch");
  DRGNN_REQUIRE(step3_dims_ok(s->kind, s->F, s->h1, s->h2, s->Hd, s->out),
                "net_step: unsupported shape (kind %d, F %d, h1 %d, h2 %d: widths must be multiples of 4)", s->kind, s->F, s->h1, s->h2);
  DRGNN_REQUIRE(s->x && s->blob && s->params && s->pred && s->status, "net_step: NULL input");
  DRGNN_REQUIRE(s->gdesc || (s->node_ptr && s->edge_ptr), "net_step: needs gdesc or node_ptr + edge_ptr");
  DRGNN_REQUIRE(s->kind != 1 || s->wblob, "net_step: sGAT needs the edge weights of the structure pass (wblob)");
  DRGNN_REQUIRE(s->kind == 0 || (s->off_b1 >= 0 && s->off_b2 >= 0), "net_step: conv biases missing");
  DRGNN_REQUIRE(s->task >= 0 && s->task <= 3, "net_step: bad task %d", s->task);
  DRGNN_REQUIRE(((uintptr_t)s->zin1 % 16) == 0, "net_step: zin1 must be 16-byte aligned");
  DRGNN_REQUIRE(!s->layers3 || (s->kind != 0 && s->off_w3 >= 0 && s->off_b3 >= 0 && !(s->flags & 1)),
                "net_step: the three-layer variant needs kind 1 / 2, conv3 offsets and no mirror flag");
  DRGNN_REQUIRE(((uintptr_t)s->x % 16) == 0 && ((uintptr_t)s->blob % 16) == 0 && ((uintptr_t)s->params % 16) == 0 &&
                    s->off_fc1w % 4 == 0 && (s->wblob == nullptr || ((uintptr_t)s->wblob % 16) == 0),
                "net_step: x / blob / params must be 16-byte aligned");
  const bool train = !s->forward_only && s->task != 0;
  if (train) {
    DRGNN_REQUIRE(s->partial && s->grads && s->n_params > 0 && s->partial_ld > s->n_params && s->partial_ld % 4 == 0 &&
                      ((uintptr_t)s->partial % 16) == 0,
                  "net_step: bad gradient buffers");
    DRGNN_REQUIRE(s->task == 3 ? (s->y_class != nullptr) : (s->y != nullptr), "net_step: missing targets");
    DRGNN_REQUIRE(!s->fuse_adam || (s->adam_p && s->adam_m && s->adam_v && s->step_dev), "net_step: fuse_adam needs the Adam buffers");
  }
  DRGNN_REQUIRE(s->keep || s->drop_p <= 0.f || s->step_dev, "net_step: hashed dropout needs step_dev");
  DRGNN_REQUIRE(!(s->flags & 1) || (s->kptr0 && s->kptr1 && s->Zin1 && s->Z1 && s->arg0 && s->Zin2 && s->Z2 && s->arg1),
                "net_step: mirror flag without the mirror buffers");
  if (s->B == 0) return DRGNN_OK;
  int tiles = s->tiles;
  if (tiles == 0) {
    tiles = drgnn_net_step_pick_tiles_l(s->kind, s->F, s->h1, s->h2, s->max_n, s->max_k, s->max_q, s->max_e, s->Hd, s->out, s->layers3);
    if (tiles < 0)
      return fail(DRGNN_ERR_UNSUPPORTED, "net_step: a graph of %d nodes / %d edges does not fit a cluster of 8 CTAs", s->max_n, s->max_e);
  }
  const int64_t smem = drgnn_net_step_smem_bytes_l(s->kind, tiles, s->F, s->h1, s->h2, s->max_n, s->max_k, s->max_q, s->max_e, s->Hd,
                                                   s->out, s->layers3);
  if (smem < 0)
    return fail(DRGNN_ERR_UNSUPPORTED, "net_step: kind %d with %d tiles does not fit (max_n %d, max_e %d)", s->kind, tiles, s->max_n, s->max_e);
  int rc = step3_configure(smem);
  if (rc) return rc;
  const int stage = step3_stage(s->kind, tiles, s->F, s->h1, s->h2, s->max_n, s->max_k, s->max_q, s->max_e, s->Hd, s->out, s->layers3);
  Step3Plan plan = step3_plan(s->kind, tiles, stage, s->F, s->h1, s->h2, s->max_n, s->max_k, s->max_q, s->max_e, s->Hd, s->out,
                              s->layers3);
  const int cs = plan.cs;
  const int occ = step3_max_clusters(cs, smem);
  plan.fused_reduce = (train && !s->skip_reduce && s->step_dev != nullptr && !(s->flags & 2) && s->B <= occ) ? 1 : 0;
  drgnn_peer_comm comm;
  memset(&comm, 0, sizeof(comm));
  if (s->comm != nullptr && train && !s->skip_reduce) {
    const drgnn_peer_comm* c = s->comm;
    DRGNN_REQUIRE(c->world >= 1 && c->world <= DRGNN_MAX_PEERS && c->rank >= 0 && c->rank < c->world,
                  "net_step: bad world / rank %d / %d", c->world, c->rank);
    if (c->world > 1) {
      if (!plan.fused_reduce)
        return fail(DRGNN_ERR_UNSUPPORTED, "net_step: the in-kernel exchange needs the co-resident grid (B %d, clusters %d)", s->B, occ);
      DRGNN_REQUIRE(s->fuse_adam, "net_step: the in-kernel exchange applies Adam (fuse_adam)");
      DRGNN_REQUIRE(c->ctr && c->max_blocks >= cs * s->B && c->stride >= s->n_params + 1, "net_step: exchange layout too small");
      for (int r = 0; r < c->world; ++r) DRGNN_REQUIRE(c->xll[r], "net_step: rank %d has no low-latency exchange buffer", r);
      comm = *c;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(cs * s->B));
  cfg.blockDim = dim3(S3_THREADS);
  cfg.dynamicSmemBytes = (size_t)smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  drgnn_net_step_args k = *s;
  k.tiles = tiles;
  if (s->F % 8 || s->h1 % 8 || s->h2 % 8) k.flags &= ~4;      // the tensor-core tiles need widths that are multiples of 8
  DRGNN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, net_graph_step3_kernel, k, plan, comm));
  g_step3_tiles = tiles;
  g_step3_launches = 1;
  if (train && !s->skip_reduce && !plan.fused_reduce) {
    net_step_reduce_kernel<<<(s->n_params + 1 + 31) / 32, 32 * RED3_SPLITS, 0, (cudaStream_t)stream>>>(k);
    DRGNN_CHECK_LAUNCH("net_step_reduce_kernel");
    g_step3_launches = 2;
  }
  return DRGNN_OK;
}

extern "C" int drgnn_sgat_step(const drgnn_net_step_args* s, void* stream) {
  DRGNN_REQUIRE(s != nullptr && s->kind == 1, "sgat_step: args must carry kind 1 (sGAT)");
  return drgnn_net_step(s, stream);
}
extern "C" int drgnn_fout_step(const drgnn_net_step_args* s, void* stream) {
  DRGNN_REQUIRE(s != nullptr && s->kind == 2, "fout_step: args must carry kind 2 (FoutNet)");
  return drgnn_net_step(s, stream);
}

extern "C" int drgnn_debug_cta_times(uint64_t* out, int32_t ctas) {
  DRGNN_REQUIRE(out != nullptr && ctas >= 0 && ctas <= 2048, "debug_cta_times: bad arguments");
  DRGNN_CHECK_CUDA(cudaMemcpyFromSymbol(out, g_cta_times, sizeof(unsigned long long) * 2 * (size_t)ctas));
  return DRGNN_OK;
}

extern "C" int drgnn_debug_phase3_cycles(uint64_t* out32) {
  DRGNN_REQUIRE(out32 != nullptr, "debug_phase3_cycles: NULL");
  DRGNN_CHECK_CUDA(cudaMemcpyFromSymbol(out32, g_phase3, sizeof(unsigned long long) * 32));
  return DRGNN_OK;
}
