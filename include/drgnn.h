/*
 * drgnn.h - C-ABI of the B200-native DeepRank-GNN hot path (libdrgnn.so).
 *
 * The reference (DeepRank/Deeprank-GNN v0.1.4) is pure Python and has no FFI of its
 * own; its native work happens inside torch_scatter / torch_sparse / torch_geometric /
 * ATen.  Each entry point below replaces one of those third-party kernels at the
 * reference call site cited next to it (paths relative to the reference root).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *    parameter name starts with h_;  float = fp32, indices = int32 unless stated
 *  - the caller owns every buffer (inputs, outputs, workspaces); the library never
 *    allocates or frees device memory and keeps no pointer after return
 *  - all work is enqueued on `stream` (a cudaStream_t passed as void*), no implicit
 *    synchronisation, no host read-back
 *  - return 0 on success, <0 on error; drgnn_last_error() gives the message of the
 *    last failing call of the calling thread
 *  - row-major [rows, cols] matrices with an explicit leading dimension (ld*) in
 *    elements, so column slices of wider buffers can be used in place
 */
#ifndef DRGNN_H
#define DRGNN_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DRGNN_OK 0
#define DRGNN_ERR_INVALID (-1)
#define DRGNN_ERR_CUDA (-2)
#define DRGNN_ERR_UNSUPPORTED (-3)

/* status bits written (OR-ed) by the structure kernels into io->status[0] */
#define DRGNN_ST_EDGE_OUTSIDE_GRAPH 1   /* an edge endpoint is not a node of its graph          */
#define DRGNN_ST_CLUSTER_RANGE 2        /* max-min cluster id of one graph exceeds the bitmap cap */
#define DRGNN_ST_CLUSTER1_LENGTH 4      /* len(cluster1 of graph g) != n_unique(cluster0 of g)   */
#define DRGNN_ST_CLUSTER_ORDER 8        /* global cluster ids are not increasing with graph id   */
#define DRGNN_ST_NEGATIVE_ID 16         /* negative cluster id                                   */
#define DRGNN_ST_FUSED_BOUNDS 64        /* a graph exceeds the max_n / max_k / max_q given to the fused kernels */

const char* drgnn_last_error(void);
int drgnn_version(void);
/* multiprocessor count / max opt-in shared memory of the current device (cached) */
int drgnn_device_sms(void);
int drgnn_device_smem_optin(void);

/* ------------------------------------------------------------------------------------
 * 1. Structure pass (integer, bit-exact): everything that depends only on the batch.
 *    Replaces, for a whole mini-batch in two launches:
 *      get_preloaded_cluster            deeprank_gnn/community_pooling.py:25-30
 *      PyG consecutive_cluster          community_pooling.py:197 (and inside max_pool_x,
 *                                       ginet.py:114,129 sGAT.py:130 foutnet.py:117)
 *      PyG pool_edge (+torch_sparse coalesce)   community_pooling.py:204-205
 *      PyG pool_batch                   community_pooling.py:222-224
 *    and builds the CSR / CSC forms the aggregation kernels consume instead of the
 *    reference's x[col] / x[row] gathers and scatter_add (ginet.py:57-71, sGAT.py:70-81,
 *    foutnet.py:71-73).
 *    One CTA per graph; graphs are contiguous node / edge ranges given by node_ptr /
 *    edge_ptr (as produced by Batch.from_data_list).
 * ---------------------------------------------------------------------------------- */
#define DRGNN_BLOB_HEADER 32
#define DRGNN_BLOB_OFFSET(g, n0, e0) (48 * (int64_t)(g) + 12 * (int64_t)(n0) + 4 * (int64_t)(e0))
#define DRGNN_BLOB_WORDS(B, N, E) (48 * (int64_t)(B) + 12 * (int64_t)(N) + 4 * (int64_t)(E) + 16)
/* words of graph g's block that carry data: header + 9n + 5 + 3m, rounded up to 4 */
#define DRGNN_BLOB_USED(n, m) ((DRGNN_BLOB_HEADER + 9 * (n) + 5 + 3 * (m) + 3) & ~3)
typedef struct drgnn_structure_io {
  /* ---- sizes ---- */
  int32_t B;            /* graphs                                                        */
  int32_t N;            /* nodes  (= node_ptr[B])                                        */
  int32_t E;            /* directed edges (= edge_ptr[B])                                */
  int32_t L1;           /* len(cluster1) (0 if absent); must equal K0 for valid data      */
  int32_t ne;           /* edge_attr width (0 if edge_attr == NULL)                       */
  int32_t max_n;        /* max nodes of one graph   (host-known from node_ptr)            */
  int32_t max_e;        /* max edges of one graph   (host-known from edge_ptr)            */
  int32_t clusters_are_local; /* 1: ids are per-graph local (un-offset, DataSet.py:348-357);
                                 0: ids are already global (must increase with graph id)  */
  int32_t idx32;        /* 0: edge_index / cluster0 / cluster1 are int64 (reference tensors);
                           1: they are int32 (packed feeder batches); 2: cluster0 / cluster1 are uint16
                           (compact feeder records; needs edge16 != 0)                     */
  int32_t edge16;       /* 1: edge_index is uint16 [2,E] holding graph-LOCAL node ids (compact feeder
                           batches: half the bytes over PCIe); cluster ids are int32 or uint16 (idx32 = 1 | 2);
                           2: edge_index is uint16 [2,E/2]: only the FIRST half of every graph's directed edges
                           travels - the loader stores each edge in both directions, first half i -> j, second
                           half j -> i (DataSet.py:266-269), so edge m/2 + e of a graph is edge e mirrored; E and
                           edge_ptr still count directed edges (all even), the graph's pairs start at edge_ptr[g]/2 */
  /* ---- inputs ---- */
  const int32_t* node_ptr;   /* [B+1] */
  const int32_t* edge_ptr;   /* [B+1] */
  const int32_t* c1_ptr;     /* [B+1] segment pointers of cluster1 (NULL if L1 == 0)       */
  const void* edge_index;    /* [2,E]  row = edge_index[0] = destination / segment id,
                                        col = edge_index[1] = gathered source (ginet.py:52) */
  const float* edge_attr;    /* [E,ne] or NULL                                            */
  const void* cluster0;      /* [N]                                                       */
  const void* cluster1;      /* [L1] or NULL                                              */
  /* ---- level-0 graph: final outputs ---- */
  int32_t* rowptr0;  /* [N+1] CSR by destination; within a row ascending original edge id */
  int32_t* col0;     /* [E]   source node of CSR slot                                     */
  int32_t* eid0;     /* [E]   original edge id of CSR slot                                */
  int32_t* cscptr0;  /* [N+1] CSR of the transposed graph (by source)                     */
  int32_t* cscrow0;  /* [E]   destination node of CSC slot                                */
  int32_t* csceid0;  /* [E]   original edge id of CSC slot                                */
  float* w0csr;      /* [E]   edge_attr[eid0[p],0]    (NULL to skip; needs ne >= 1)       */
  float* w0csc;      /* [E]   edge_attr[csceid0[p],0] (NULL to skip)                      */
  /* ---- level-0 clustering ---- */
  int32_t* cl0;      /* [N]   dense pooled-node id of every node (consecutive_cluster inv) */
  int64_t* cl0_i64;  /* [N]   same as int64 (API mirror; NULL to skip)                    */
  int32_t* cmptr0;   /* [N+1] members-of-cluster CSR: pointers (first K0+1 entries valid)  */
  int32_t* cmem0;    /* [N]   member node ids, ascending inside a cluster                 */
  int32_t* kptr0;    /* [B+1] pooled-node range of every graph                            */
  int32_t* batch1;   /* [N]   graph id of pooled node (first K0 valid)  (pool_batch)       */
  int64_t* batch1_i64;/* [N]  API mirror, NULL to skip                                     */
  /* ---- level-1 (pooled) graph ---- */
  int32_t* rowptr1;  /* [N+1] (first K0+1 valid) CSR of the coarsened graph               */
  int32_t* col1;     /* [E]   (first E1 valid)                                            */
  int64_t* edge_index1; /* [2,E] int64 mirror laid out with row stride E (first E1 columns
                           valid): sorted by (row,col), unique, no self loops; NULL to skip */
  float* edge_attr1; /* [E,ne] summed attrs of merged edges (first E1 rows); NULL to skip  */
  int32_t* cscptr1;  /* [N+1] */
  int32_t* cscrow1;  /* [E]   */
  int32_t* csceid1;  /* [E]   pooled-edge id (position in col1) of CSC slot                */
  float* w1csc;      /* [E]   edge_attr1[csceid1[p],0] (NULL to skip)                      */
  /* ---- level-1 clustering (only if cluster1 != NULL) ---- */
  int32_t* cl1;      /* [L1]  dense id of pooled node -> second-level cluster             */
  int32_t* cmptr1;   /* [L1+1] */
  int32_t* cmem1;    /* [L1]  */
  int32_t* kptr1;    /* [B+1] second-level node range per graph (readout segments)        */
  int32_t* batch2;   /* [L1]  (first K1 valid) */
  int64_t* batch2_i64;
  /* ---- counts / status ---- */
  int32_t* counts;   /* [4]: K0, E1, K1, reserved                                         */
  int32_t* status;   /* [1]: OR of DRGNN_ST_* bits (caller zeroes before the call)        */
  /* ---- workspaces ---- */
  int32_t* gstat;    /* [8*B] per-graph counts                                            */
  int32_t* scratch_n;/* [6*(N+B)+L1+8] node-indexed scratch                               */
  int32_t* scratch_e;/* [4*E+8] edge-indexed scratch                                      */
  float* scratch_f;  /* [E*max(ne,1)] pooled attr scratch                                 */
  /* ---- per-graph structure blob (optional, NULL to skip) ----
   * Everything integer the per-graph fused kernels need of graph g, with graph-LOCAL indices, in
   * one contiguous 16-byte aligned block so that a CTA stages it with ONE bulk copy and depends on
   * nothing the cross-graph finalize kernel computes.  Block g starts at word
   * DRGNN_BLOB_OFFSET(g, node_ptr[g], edge_ptr[g]); layout (n nodes, m directed edges):
   *   header[32]: [0] n, [1] m, [2] K0, [3] E1, [4] K1, [5] 1 when complete, rest 0
   *   rowptr0[n+1] col0[m] | rowptr1[n+1] col1[m] | cmptr0[n+1] cmem0[n] cl0[n] |
   *   cmptr1[n+1] cmem1[n] cl1[n] | cscptr1[n+1] cscrow1[m]      (capacity by n / m; K0+1, E1, ...
   *   entries are valid).  Size of the buffer: DRGNN_BLOB_WORDS(B, N, E) int32. */
  int32_t* blob;
  /* Edge weights of the blob's neighbour lists (optional, NULL to skip; needs edge_attr, uses column 0):
   * a float array PARALLEL to blob (same size, same per-graph offsets) that holds, at the word offsets of
   * col0 / col1 / cscrow1 inside graph g's block, edge_attr of the level-0 CSR slot, the summed attribute of
   * the pooled edge (coalesce, community_pooling.py:204-205) in pooled-CSR order and in pooled-CSC order.
   * What the fused sGAT step (drgnn_net_step, kind 1) stages next to the index lists (sGAT.py:76). */
  float* wblob;
  /* First aggregation of the network, computed next to the structure (optional, zin1 == NULL to skip; honoured by
   * drgnn_structure_blob only).  The input rows of conv1's dense transform depend on the batch alone, so the
   * structure pass - which runs on a side stream while the previous step computes - can leave them ready for the
   * step kernels (drgnn_ginet_step_args.zin1 / drgnn_net_step_args.zin1), which then stage them instead of the feature tile:
   *   zin_kind 0 (GINet, ginet.py:57-71):    zin1[i] = sum_e x[col]                                     (ld_zin1 >= F)
   *   zin_kind 1 (sGAT, sGAT.py:70-92):      zin1[i] = [ s_i x_i | mean_e a_e x_col | 1 0 0 0 ]         (ld_zin1 >= 2F + 4)
   *   zin_kind 2 (FoutNet, foutnet.py:62-80): zin1[i] = [ x_i | (1/deg) sum_e x_col | 1 0 0 0 ], deg 0 -> NaN
   * rows of node i at zin1 + i * ld_zin1, summed in CSR-slot order (bit-identical to the step kernels' own phase).
   * x [N, F] fp32 row-major, F % 4 == 0, ld_zin1 % 4 == 0. */
  const float* x;
  float* zin1;
  int32_t F; int32_t ld_zin1; int32_t zin_kind;
  /* reserved (0) */
  int32_t launch_flags;
  /* max_k / max_q (drgnn_structure_blob; 0 = unknown -> max_n): host bounds of the level-0 / level-1 cluster counts
   * of ONE graph.  The pass sizes its pooled-graph bitmaps by them (shared memory: K x K/32 words instead of
   * n x n/32), which lets several CTAs share an SM and lets graphs of a thousand nodes take this pass; a graph that
   * exceeds them is flagged (DRGNN_ST_FUSED_BOUNDS) and its blob left incomplete. */
  int32_t max_k; int32_t max_q;
} drgnn_structure_io;

/* Dynamic shared memory the per-graph kernel needs for (max_n, max_e); <0 if a graph is
 * too large for one CTA (DRGNN_ERR_UNSUPPORTED). */
int64_t drgnn_structure_smem_bytes(int32_t max_n, int32_t max_e, int32_t max_c1);
/* ... of drgnn_structure_blob for graphs of up to max_n nodes / max_e directed edges / max_k and max_q clusters of the
 * two levels (0: max_n), with (weights != 0) or without the sGAT edge weights, with a feature tile of x_words floats
 * staged for the first aggregation (0: none); <0 when it does not fit */
int64_t drgnn_structure_blob_smem_bytes_ex(int32_t max_n, int32_t max_e, int32_t max_k, int32_t max_q, int32_t weights,
                                           int32_t x_words);
int drgnn_structure_build(const drgnn_structure_io* io, void* stream);

/* get_preloaded_cluster as a stand-alone op (community_pooling.py:25-30):
 * cluster[i] += sum_{g < graph(i)} (max(cluster of g) + 1), in place, int64.
 * seg_ptr [B+1] gives the contiguous segment of every graph. work: [B] int64. */
int drgnn_cluster_offset(int64_t* cluster, const int32_t* seg_ptr, int32_t B, int64_t* work, void* stream);

/* segment pointers from a sorted int64 id vector (`batch`): ptr[g] = first i with ids[i] >= g.
 * status bit0 is set if ids is not sorted ascending or holds a value outside [0,B). */
int drgnn_ptr_from_sorted_ids(const int64_t* ids, int32_t n, int32_t B, int32_t* ptr, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------
 * 2. Aggregation: the edge gather -> (weight) -> segmented reduction that replaces
 *    x[col] + scatter_sum (ginet.py:57-71), scatter_mean (sGAT.py:70-81) and the per-node
 *    Python loop of FoutLayer (foutnet.py:71-73); the same kernel on the CSC form is the
 *    backward.  No atomics: one sub-warp per destination row, deterministic.
 *
 *    out[i, 0:C] = act( selfc_i * self_src[i, 0:C]
 *                       + post_i * sum_{p in [rowptr[i], rowptr[i+1])} ew[p] * sscale[col[p]] * src[col[p], 0:C]
 *                       + bias[0:C] )
 *    post_mode : 0 -> post_i = 1 (sum) ; 1 -> 1/max(deg_i,1) (scatter_mean) ;
 *                2 -> 1/deg_i, deg_i = 0 gives NaN (torch.mean of an empty set, foutnet.py:73)
 *    self_mode : 0 -> no self term ; 1 -> selfc_i = 1 ; 2 -> selfc_i = post_i * sum_p ew[p]
 *                (also stored to selfc_out[i] when selfc_out != NULL) ; 3 -> selfc_i = selfc_in[i]
 *    ew, sscale, bias, self_src may be NULL.  relu != 0 applies max(.,0).
 *    self_out != NULL additionally stores selfc_i * self_src[i,:] there (ld = ld_self_out)
 *    INSTEAD of adding the self term to out.
 *    post_out != NULL stores post_i (the backward on the CSC form uses it as sscale).
 *    n_rows_dev (may be NULL) holds the live row count on the device (<= n_rows).
 * ---------------------------------------------------------------------------------- */
typedef struct drgnn_aggregate_args {
  const float* src; int32_t ld_src;
  float* out; int32_t ld_out;
  const int32_t* rowptr; const int32_t* col;
  const float* ew; const float* sscale;
  const float* self_src; int32_t ld_self;
  float* self_out; int32_t ld_self_out;
  const float* selfc_in; float* selfc_out;
  float* post_out;
  const float* bias;
  int32_t n_rows; const int32_t* n_rows_dev;
  int32_t C; int32_t post_mode; int32_t self_mode; int32_t relu;
} drgnn_aggregate_args;
int drgnn_aggregate(const drgnn_aggregate_args* a, void* stream);

/* Same contract, per-graph tiles: everything one graph (tile) needs - its source rows
 * tile_ptr[t]..tile_ptr[t+1], its rowptr slice and its col / ew slices tile_eptr[t]..tile_eptr[t+1]
 * (tile_eptr[t] == rowptr[tile_ptr[t]]) - is streamed into shared memory with bulk async copies
 * (cp.async.bulk + mbarrier), double buffered by persistent CTAs, so neighbour gathers and
 * index reads never leave the SM and HBM traffic is compulsory-only.
 * Requires every col[p] of a tile to lie inside the tile, ld_src == C, and rowptr / col / ew to
 * be 16-byte aligned and READABLE up to the next multiple of 4 elements past their end. */
int drgnn_aggregate_tiled(const drgnn_aggregate_args* a, const int32_t* tile_ptr, const int32_t* tile_eptr,
                          int32_t n_tiles, int32_t max_tile_rows, int32_t max_tile_edges, void* stream);

/* ------------------------------------------------------------------------------------
 * 3. Dense per-node transform (the nn.Linear / torch.mm of ginet.py:57-58,137-139,
 *    sGAT.py:73,134-135, foutnet.py:62,65,121-122), done on N rows instead of E rows.
 *    Y[r, g*Fout + o] = act( sum_k X[r, g*Fin + k] * Wg[k,o] + bias[g*Fout+o] )
 *    w_layout 0: W stored [groups][Fout][Fin] (nn.Linear.weight) ; 1: [groups][Fin][Fout]
 *    (sGAT weight / Fout Wc,Wn / transposed use in backward).
 *    out_mask != NULL: Y is multiplied by (out_mask[r,c] > 0) * mask_scale after the
 *    activation (out_mask has Y's shape, leading dimension ld_mask).  One mechanism, two uses:
 *    train-mode dropout (ginet.py:138; out_mask = 0/1 keep mask, mask_scale = 1/(1-p)) and
 *    the fused ReLU / dropout backward (out_mask = the forward activation).
 *    math: 0 = fp32 FMA, 1 = 3xTF32 on mma.sync tensor-core tiles (error-compensated, ~fp32 accuracy),
 *          2 = 3xTF32 on the 5th-generation tensor cores: tcgen05.mma.kind::tf32, M = 128 row tiles, fp32
 *              accumulator in TMEM, operands in the canonical K-major shared-memory layout (csrc/linear_tc5.cu);
 *              shapes it does not take (drgnn_linear_tcgen05_supported == 0: groups > 1, Fin % 8, Fin > 64,
 *              Fout other than 16 / 32 / 64, unaligned rows) fall back to mode 1
 * ---------------------------------------------------------------------------------- */
typedef struct drgnn_linear_args {
  const float* X; int32_t ldx;
  const float* W; const float* bias;
  float* Y; int32_t ldy;
  const float* out_mask; int32_t ld_mask; float mask_scale;
  int32_t rows; const int32_t* rows_dev;
  int32_t Fin; int32_t Fout; int32_t groups;
  int32_t w_layout; int32_t relu; int32_t math;
} drgnn_linear_args;
int drgnn_linear(const drgnn_linear_args* a, void* stream);
int drgnn_linear_tcgen05_supported(const drgnn_linear_args* a);
int drgnn_linear_tcgen05(const drgnn_linear_args* a, void* stream);
/* diagnostic: cycles CTA 0 of the last tcgen05 launch spent per phase, summed over its tiles: [0] TF32 split + operand
 * stores, [1] fence + barrier + MMA issue, [2] prefetch issue, [3] wait for the MMAs, [4] TMEM read-back + output
 * stores, [5] closing barrier, [6] tiles.  Synchronises the device. */
int drgnn_debug_tc5_cycles(uint64_t* out8);

/* Weight / bias gradient of the transform: dW[g][o][k] (w_layout 0) or dW[g][k][o]
 * (w_layout 1) (+)= sum_r G[r, g*Fout+o] * X[r, g*Fin+k], dbias[g*Fout+o] (+)= sum_r G[r,..].
 * Deterministic reduction over row chunks: partials in `work`, summed in a fixed order (small
 * matrices: one launch, the last CTA to finish does the sum; large ones: a second launch).
 * `work` must be ZERO-FILLED before its first use (its last 4 floats hold a ticket counter that
 * the kernel resets itself) and must not be shared by launches that may run concurrently.
 * accumulate != 0 adds into dW / dbias, else overwrites.  dbias may be NULL. */
typedef struct drgnn_linear_wgrad_args {
  const float* X; int32_t ldx;
  const float* G; int32_t ldg;
  float* dW; float* dbias;
  int32_t rows; const int32_t* rows_dev;
  int32_t Fin; int32_t Fout; int32_t groups;
  int32_t w_layout; int32_t accumulate;
  float* work; int64_t work_floats;
} drgnn_linear_wgrad_args;
int64_t drgnn_linear_wgrad_work_floats(int32_t rows, int32_t Fin, int32_t Fout, int32_t groups);
int drgnn_linear_wgrad(const drgnn_linear_wgrad_args* a, void* stream);

/* ------------------------------------------------------------------------------------
 * 4. Cluster max-pool (torch_scatter.scatter_max, community_pooling.py:201; PyG max_pool_x)
 *    y[k,c] = max_{i in members(k)} x[i,c]; argmax[k,c] = FIRST member attaining it.
 *    Backward routes the gradient to argmax only:
 *    dx[i,c] = (argmax[cl[i],c] == i) ? g[cl[i],c] : 0, optionally times (relu_out[i,c] > 0).
 * ---------------------------------------------------------------------------------- */
int drgnn_maxpool_fwd(const float* x, int32_t ldx, const int32_t* cmptr, const int32_t* cmem,
                      int32_t n_clusters, const int32_t* n_clusters_dev, int32_t C,
                      float* y, int32_t ldy, int32_t* argmax, void* stream);
int drgnn_maxpool_bwd(const float* g, int32_t ldg, const int32_t* argmax, const int32_t* cl,
                      const float* relu_out, int32_t ld_relu, int32_t n_nodes,
                      const int32_t* n_nodes_dev, int32_t C, float* dx, int32_t lddx, void* stream);

/* ------------------------------------------------------------------------------------
 * 5. Graph read-out (torch_scatter.scatter_mean(x, batch), ginet.py:133-134, sGAT.py:133,
 *    foutnet.py:120): r[b,:] = mean of rows seg_ptr[b]..seg_ptr[b+1] (0 for empty).
 *    Backward: dx[k,:] = g[b(k),:] / max(count_b,1).
 * ---------------------------------------------------------------------------------- */
int drgnn_segment_mean_fwd(const float* x, int32_t ldx, const int32_t* seg_ptr, int32_t B, int32_t C,
                           float* r, int32_t ldr, void* stream);
int drgnn_segment_mean_bwd(const float* g, int32_t ldg, const int32_t* seg_ptr, int32_t B, int32_t C,
                           float* dx, int32_t lddx, void* stream);

/* ------------------------------------------------------------------------------------
 * 6. Loss, optimiser (NeuralNet.py:239-263, 500-503): fused loss + dLoss/dpred, flat Adam.
 *    mse:  loss = mean_b (pred_b - y_b)^2 over B_global ;   dpred = 2 (pred - y) / B_global
 *    ce :  weighted CrossEntropyLoss(reduction='mean'); dlogits accordingly.
 *    loss_out[0] receives the LOCAL contribution sum_b(...)/B_global (all-reduce it to get
 *    the global mean when the batch is sharded over ranks).
 * ---------------------------------------------------------------------------------- */
int drgnn_mse_loss(const float* pred, const float* y, int32_t B_local, float inv_B_global,
                   int32_t sigmoid, float* loss_out, float* dpred, void* stream);
int drgnn_ce_loss(const float* logits, int32_t ld, const int64_t* target, const float* class_w,
                  int32_t B_local, int32_t n_classes, float inv_norm_global, float* loss_out,
                  float* dlogits, void* stream);
/* torch.optim.Adam (no weight decay, no amsgrad): step_dev[0] holds the step count (float,
 * incremented on the device so the call is CUDA-graph replayable). grad_scale multiplies g. */
int drgnn_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                    float* step_dev, int64_t n, float lr, float beta1, float beta2, float eps,
                    float grad_scale, void* stream);

/* ------------------------------------------------------------------------------------
 * 7. Fused network head (ginet.py:136-139, sGAT.py:134-135, foutnet.py:121-122 + the loss of
 *    NeuralNet.py:500 + the backward of both): ONE launch for
 *      H = relu(R W1^T + b1) [* keep * keep_scale],  pred = H W2^T + b2,
 *      loss / dLoss/dpred (task 1: MSE, 2: MSE of sigmoid(pred), 3: class-weighted cross entropy,
 *      0: forward only), dW2, db2, dW1, db1 (overwritten), dR = dLoss/dR.
 *    W1 [Hd, C] and W2 [out, Hd] are nn.Linear weights.  keep: [B, Hd] 0/1 floats or NULL (then
 *    keep_scale must be 1).  Gradient pointers may all be NULL (no backward).  H (optional):
 *    [B, Hd] copy of the hidden activation.  Meant for B up to a few hundred rows (one CTA).
 * ---------------------------------------------------------------------------------- */
typedef struct drgnn_head_args {
  const float* R; int32_t ldr;
  const float* W1; const float* b1;
  const float* W2; const float* b2;
  const float* keep; float keep_scale;
  const float* y; const int64_t* y_class; const float* class_w;
  int32_t B; int32_t C; int32_t Hd; int32_t out;
  int32_t task; float inv_norm;
  float* pred; float* loss; float* H;
  float* dW1; float* db1; float* dW2; float* db2;
  float* dR; int32_t lddr;
} drgnn_head_args;
int64_t drgnn_head_smem_bytes(int32_t C, int32_t Hd, int32_t out);   /* <0: does not fit one CTA */
int drgnn_head(const drgnn_head_args* a, void* stream);

/* ------------------------------------------------------------------------------------
 * 8. Per-graph fused GINet (ginet.py:99-134 and its autograd backward): ONE CTA per graph keeps
 *    every intermediate of the graph in shared memory.
 *      fwd: x -> AX = A x -> Z1 = relu(AX W1^T) -> cluster max -> AP = A1 P1 -> Z2 = relu(AP W2^T)
 *           -> level-1 cluster max -> R = per-graph mean           (both branches: W1 [nb*h1, F],
 *           W2 [nb][h2][h1]); stores Zin1 = AX [N,F], Z1 [N,nb*h1], arg0, Zin2 = AP, Z2, arg1, R.
 *      bwd: dR [B, nb*h2] -> dW1 [nb*h1, F], dW2 [nb][h2][h1] (per-graph partials in `partial`
 *           [B, nb*h1*F + nb*h2*h1], summed in graph order: deterministic).
 *    Structure arrays are the outputs of drgnn_structure_build.  max_n / max_k / max_q: host-known
 *    upper bounds of nodes, level-0 clusters and level-1 clusters of ONE graph (shared-memory
 *    sizing; a violation sets DRGNN_ST_FUSED_BOUNDS in status[0]).  Needs F % 4 == 0, h1 % 4 == 0,
 *    h2 % 8 == 0, (nb*h1) % 8 == 0.
 * ---------------------------------------------------------------------------------- */
typedef struct drgnn_ginet_fused_args {
  int32_t B; int32_t F; int32_t h1; int32_t h2; int32_t nb;
  int32_t max_n; int32_t max_k; int32_t max_q;
  const int32_t* node_ptr;
  const int32_t* rowptr0; const int32_t* col0;
  const int32_t* rowptr1; const int32_t* col1;
  const int32_t* cscptr1; const int32_t* cscrow1;
  const int32_t* cmptr0; const int32_t* cmem0; const int32_t* cl0; const int32_t* kptr0;
  const int32_t* cmptr1; const int32_t* cmem1; const int32_t* cl1; const int32_t* kptr1;
  int32_t* status;
  const float* W1; const float* W2;
  const float* x;
  float* Zin1; float* Z1; int32_t* arg0;
  float* Zin2; float* Z2; int32_t* arg1;
  float* R;
  const float* dR; float* partial; float* dW1; float* dW2;
} drgnn_ginet_fused_args;
int64_t drgnn_ginet_fused_smem_bytes(int32_t F, int32_t h1, int32_t h2, int32_t nb, int32_t max_n, int32_t max_k,
                                     int32_t max_q, int32_t backward);   /* <0: does not fit one CTA */
int drgnn_ginet_fused_fwd(const drgnn_ginet_fused_args* a, void* stream);
int drgnn_ginet_fused_bwd(const drgnn_ginet_fused_args* a, void* stream);

/* Whole GINet training step of every graph in ONE launch (+ one reduction launch): forward as
 * above, the network head on the graph's own read-out row (fc1 [Hd, nb*h2] / ReLU / dropout keep
 * mask / fc2 [out, Hd] are row-wise, the loss is a sum of per-graph terms: task 1 MSE, 2 MSE of
 * sigmoid, 3 class-weighted cross entropy, see drgnn_head), and the backward down to per-graph
 * partials of EVERY parameter gradient.  partial [B, partial_ld] rows are laid out like the flat
 * gradient buffer `grads` [n_params] (offsets off_*; slot n_params of a row holds the graph's loss
 * term); rows must be ZERO in the slots no live parameter owns.  The reduction sums the rows in
 * graph order (deterministic) into grads and loss.  forward_only != 0: predictions only. */
struct drgnn_peer_comm;
typedef struct drgnn_ginet_step_args {
  drgnn_ginet_fused_args g;            /* dR / partial / dW1 / dW2 of the embedded block are unused */
  const float* fc1_w; const float* fc1_b; const float* fc2_w; const float* fc2_b;
  int32_t Hd; int32_t out;
  const float* keep; float keep_scale;
  const float* y; const int64_t* y_class; const float* class_w;
  int32_t task; float inv_norm;
  float* pred; float* loss;
  float* partial; int64_t partial_ld;
  float* grads; int32_t n_params;
  int32_t off_w1; int32_t off_w2; int32_t off_fc1w; int32_t off_fc1b; int32_t off_fc2w; int32_t off_fc2b;
  int32_t forward_only;
  int32_t head_off;                    /* set by the library */
  /* dropout without a mask tensor: keep == NULL and drop_p > 0 -> unit j of graph g is kept iff
   * hash(seed, (uint)step_dev[0], g*Hd+j) >= drop_p (counter-based, replayable from a CUDA graph) */
  float drop_p; uint32_t seed;
  /* fuse_adam != 0: the reduction launch also applies torch.optim.Adam to adam_p (m, v in adam_m /
   * adam_v) and increments step_dev[0]; step_dev is [4] floats ([1] is a ticket counter, zero it once) */
  int32_t fuse_adam; float lr; float beta1; float beta2; float eps;
  float* adam_p; float* adam_m; float* adam_v; float* step_dev;
  /* skip_reduce != 0: stop after the per-graph launch (partial rows written); the caller reduces
   * them itself, e.g. with drgnn_peer_reduce_adam on several GPUs */
  int32_t skip_reduce;
  /* flags bit 0: the cluster kernel also mirrors the intermediates (Zin1, Z1, arg0, Zin2, Z2, arg1)
   * to global memory (it keeps them in shared memory; the single-CTA kernel always stores them);
   * bit 1: never fuse the gradient reduction into the cluster kernel (step_dev must be [4] floats,
   * zero-initialised: [2] is the grid-barrier counter of the fused reduction);
   * bit 2 (cluster kernel): the dense products run on tensor-core tiles (mma.sync.m16n8k8 TF32 with the 3-product
   * error compensation, ~1e-6; widths must be multiples of 8, else the fp32 FMA tiles are used);
   * bit 3: block 0 records its phase clocks (drgnn_debug_phase_cycles; diagnostic, slows the launch slightly);
   * bit 6: head v2 of the cluster kernel - fc2 / loss term / dLoss/dpred evaluated by every warp (regression, out = 1:
   * no barrier and no single-thread section between fc2 and the head backward), the read-out row stored after the
   * cluster barrier of its exchange, and - with the in-kernel reduction - the fc1.weight gradient rows (77 % of a
   * per-graph gradient row) NOT stored: the reducing CTA forms dh_g[j] * R_g[c] from the fc1.bias gradient slots of
   * `partial` and the read-out rows R (rounded product, same ordered sums: bit-identical gradients); the fc1.weight
   * slots of `partial` are then unspecified after the call */
  int32_t flags;
  /* max_e: host bound of the directed edges of one graph (> 0 enables the cluster kernel: a pair of
   * CTAs per graph, one GINet branch each, nb == 2).  variant: 0 = pick (cluster kernel when it
   * fits shared memory, else single CTA), 1 = single-CTA kernel, 2 = cluster kernel or error. */
  int32_t max_e; int32_t variant;
  /* cluster kernel inputs: the per-graph structure blobs of the structure pass
   * (drgnn_structure_io.blob) and the edge pointers [B+1] of the batch; NULL -> single-CTA kernel */
  const int32_t* blob; const int32_t* edge_ptr;
  /* comm != NULL (host pointer to a drgnn_peer_comm, world > 1): the cluster kernel also runs the
   * gradient exchange of drgnn_peer_reduce_adam itself - after its grid barrier every CTA sums its
   * slice of the per-graph rows, stores it into every rank's exchange buffer over NVLink peer
   * memory, releases its flag, waits for the same CTA of every peer, sums the slots in rank order
   * and applies Adam.  Every rank must launch the same grid (equal graphs per rank); needs
   * comm->max_blocks >= 2B, the grid co-resident (drgnn_ginet_step2_max_clusters) and fuse_adam. */
  const struct drgnn_peer_comm* comm;
  /* gdesc (optional): io->gstat of a blob-only structure pass (drgnn_structure_blob) - 8 ints per graph
   * [K0, E1, K1, node_ptr[g], edge_ptr[g], m, 0, n]: the cluster kernel reads a graph's extents from this
   * L2-resident record (written a few microseconds earlier) instead of the cold node_ptr / edge_ptr */
  const int32_t* gdesc;
  /* zin1 (optional, cluster kernel): AX = A x of every node, precomputed by the structure pass
   * (drgnn_structure_io.zin1 with zin_kind 0 and ld_zin1 = F + 4): the kernel stages the graph's rows with ONE bulk
   * copy instead of the feature tile and skips its own aggregation phase (bit-identical rows) */
  const float* zin1;
} drgnn_ginet_step_args;
int64_t drgnn_ginet_step_smem_bytes(int32_t F, int32_t h1, int32_t h2, int32_t nb, int32_t max_n, int32_t max_k,
                                    int32_t max_q, int32_t Hd, int32_t out);
int drgnn_ginet_step(const drgnn_ginet_step_args* s, void* stream);
/* shared memory of one CTA of the cluster kernel (<0: unsupported shape / does not fit) and the
 * variant (1 / 2) the last drgnn_ginet_step call of this thread launched (0: none yet) */
int64_t drgnn_ginet_step2_smem_bytes(int32_t F, int32_t h1, int32_t h2, int32_t max_n, int32_t max_k, int32_t max_q,
                                     int32_t max_e, int32_t Hd, int32_t out);
int drgnn_ginet_step_last_variant(void);
/* kernels the last drgnn_ginet_step call of this thread launched: 1 when the cluster kernel also
 * reduced the gradients (+ Adam) behind a grid barrier (the whole grid co-resident: B <= clusters the
 * device can hold; flags bit 1 disables it), else 2 (per-graph kernel + reduction), 1 for scoring */
int drgnn_ginet_step_last_launches(void);
/* 2-CTA clusters of the cluster kernel the device holds at once with `smem_bytes` per CTA (<0: error) */
int drgnn_ginet_step2_max_clusters(int64_t smem_bytes);
/* diagnostic: SM clock (clock64) at the phase boundaries of the CTA that ran graph 0 in the last
 * per-graph launch: [0] start, [1] staged, [2] AX, [3] Z1, [4] P1, [5] AP, [6] Z2, [7] P2, [8] R+fc1,
 * [9] fc2, [10] loss, [11] head backward, [12] dZ2 staged, [13] dW2/dAP, [14] dP1, [15] dZ1 staged,
 * [16] dW1 (end).  Synchronises the device. */
int drgnn_debug_phase_cycles(uint64_t* out32);
/* Blob-only structure pass: ONE launch (bitmap kernel, one CTA per graph) that writes io->blob,
 * io->status and io->gstat[8g + 0..7] = [K0, E1, K1, node_ptr[g], edge_ptr[g], m, 0, n] and nothing else - what the cluster step kernel of
 * drgnn_ginet_step needs.  Replaces get_preloaded_cluster / consecutive_cluster / pool_edge +
 * coalesce / the CSR build for graphs whose bitmaps fit shared memory
 * (drgnn_structure_blob_smem_bytes >= 0); larger graphs: drgnn_structure_build (which writes the
 * blob too).  Both cluster levels are required.  status is NOT zeroed by the call. */
int64_t drgnn_structure_blob_smem_bytes(int32_t max_n, int32_t max_e);
int drgnn_structure_blob(const drgnn_structure_io* io, void* stream);
/* diagnostic: clock64 at the section boundaries of graph_blob_kernel, CTA of graph 0:
 * [0] start, [1] loaded + id extremes, [2] relabelled, [3] scattered, [4] counted + scanned, [5] end */
int drgnn_debug_blob_cycles(uint64_t* out16);
/* same for the structure pass (graph_local_kernel): [0] start, [1] edge list, [2] CSR, [3] CSC,
 * [4] relabel, [5] members, [6] coarsened edges, [7] coarsened CSC, [8] level-1 clustering (end) */
int drgnn_debug_structure_cycles(uint64_t* out32);

/* ------------------------------------------------------------------------------------
 * 8b. Whole training / scoring step of every graph of a mini-batch in ONE launch, for the three reference
 *     networks and for graphs of any size a thread-block cluster holds (csrc/fused_step3.cuh):
 *       kind 0  GINet    ginet.py:99-141     two branches on two CTA groups of the cluster
 *       kind 1  sGAT     sGAT.py:62-93, 114-138
 *       kind 2  FoutNet  foutnet.py:56-82, 103-125  (a node without neighbour gives a NaN row, foutnet.py:73)
 *     conv1 -> ReLU -> cluster max (community_pooling) -> conv2 on the coarsened graph -> ReLU -> level-1
 *     cluster max (max_pool_x) -> per-graph mean -> fc1 / ReLU / [dropout] / fc2 -> loss -> the autograd
 *     backward of all of it -> per-graph gradient rows -> (grid co-resident) ordered sum + Adam (+ the peer
 *     exchange of drgnn_peer_reduce_adam) behind a grid barrier, else a second launch.
 *     `tiles` CTAs share the nodes of one graph (rows of every level split evenly, neighbours / cluster members
 *     of other tiles read through distributed shared memory); cluster size = tiles * (kind == 0 ? 2 : 1) <= 8.
 *     Inputs: the per-graph structure blobs of the structure pass (drgnn_structure_io.blob, + wblob for kind 1),
 *     node_ptr / edge_ptr [B+1] (or gdesc = io->gstat of drgnn_structure_blob), x [N,F], the FLAT parameter
 *     buffer `params` with the offsets (in floats) of the tensors inside it:
 *       kind 0: off_w1 -> conv1.fc.weight | conv1_ext.fc.weight ([2][h1][F]), off_w2 -> [2][h2][h1], no biases
 *       kind 1: off_w1 -> conv1.weight [2F][h1], off_b1, off_w2 -> conv2.weight [2h1][h2], off_b2
 *       kind 2: off_w1 -> conv1.Wc | conv1.Wn ([2F][h1]), off_b1, off_w2 -> conv2.Wc | conv2.Wn, off_b2
 *       all  : off_fc1w [Hd][nbr*h2], off_fc1b, off_fc2w [out][Hd], off_fc2b
 *     partial [B, partial_ld] rows and grads [n_params] use the same offsets; slot n_params of a row holds the
 *     graph's loss term; rows must be ZERO in slots no live parameter owns.  task / inv_norm / keep / drop_p /
 *     fuse_adam / skip_reduce / comm / step_dev as in drgnn_ginet_step_args.  flags bit 0: mirror the
 *     intermediates to Zin1 [N,Kin1] Z1 [N,nbr*h1] arg0 Zin2 Z2 arg1 (needs kptr0 / kptr1 of
 *     drgnn_structure_build); bit 1: never fuse the gradient reduction; bit 2: the dense products (conv transforms,
 *     their input and weight gradients) run on tensor-core tiles (mma.sync.m16n8k8 TF32, 3-product error compensation,
 *     ~1e-6) instead of fp32 FMA register tiles; bit 3: block 0 records its phase clocks (drgnn_debug_phase3_cycles);
 *     bit 5: cluster c runs the graph of size rank c (largest first: shortens the tail of a grid larger than the device).
 *     R (optional): read-out rows [B, nbr*h2].
 * ---------------------------------------------------------------------------------- */
typedef struct drgnn_net_step_args {
  int32_t kind; int32_t B; int32_t F; int32_t h1; int32_t h2; int32_t Hd; int32_t out;
  int32_t max_n; int32_t max_e; int32_t max_k; int32_t max_q;
  int32_t tiles;                       /* 0: the smallest count whose shared memory fits */
  const float* x;
  const int32_t* blob; const float* wblob; const int32_t* gdesc;
  const int32_t* node_ptr; const int32_t* edge_ptr;
  const float* params;
  int32_t off_w1; int32_t off_b1; int32_t off_w2; int32_t off_b2;
  int32_t off_fc1w; int32_t off_fc1b; int32_t off_fc2w; int32_t off_fc2b;
  const float* keep; float keep_scale; float drop_p; uint32_t seed;
  const float* y; const int64_t* y_class; const float* class_w;
  int32_t task; float inv_norm; int32_t forward_only; int32_t skip_reduce;
  float* pred; float* loss; float* R;
  float* partial; int64_t partial_ld;
  float* grads; int32_t n_params;
  int32_t fuse_adam; float lr; float beta1; float beta2; float eps; int32_t flags;
  float* adam_p; float* adam_m; float* adam_v; float* step_dev;
  int32_t* status;
  const struct drgnn_peer_comm* comm;
  /* test mirrors (flags bit 0) */
  const int32_t* kptr0; const int32_t* kptr1;
  float* Zin1; float* Z1; int32_t* arg0; float* Zin2; float* Z2; int32_t* arg1;
  /* layers3 != 0 (kinds 1, 2): a THIRD conv layer h2 -> h2 on the coarsened graph between conv2 and the level-1
   * max-pool - the "sGAT 3-layer" throughput variant of BASELINE config 3 (the reference nets have two);
   * off_w3 -> conv3.weight [2h2][h2] (kind 2: conv3.Wc | conv3.Wn), off_b3 -> conv3.bias [h2] */
  int32_t layers3; int32_t off_w3; int32_t off_b3; int32_t reserved3;
  /* zin1 (optional): the input rows of conv1's transform, precomputed by the structure pass
   * (drgnn_structure_io.zin1 with zin_kind = kind and ld_zin1 = Kin1 + 4, Kin1 = F | 2F): every CTA stages its rows
   * with ONE bulk copy and skips the level-0 aggregation phase and the feature tile (bit-identical rows) */
  const float* zin1;
} drgnn_net_step_args;
/* shared memory of one CTA (<0: unsupported shape / does not fit) */
int64_t drgnn_net_step_smem_bytes(int32_t kind, int32_t tiles, int32_t F, int32_t h1, int32_t h2, int32_t max_n, int32_t max_k,
                                  int32_t max_q, int32_t max_e, int32_t Hd, int32_t out);
/* smallest tile count (1, 2, 4, 8; cluster <= 8 CTAs) whose plan fits shared memory; <0: none */
int drgnn_net_step_pick_tiles(int32_t kind, int32_t F, int32_t h1, int32_t h2, int32_t max_n, int32_t max_k, int32_t max_q,
                              int32_t max_e, int32_t Hd, int32_t out);
/* clusters of the step kernel the device holds at once for this plan (<0: error); the gradient reduction
 * (+ Adam, + peer exchange) runs inside the launch when B <= this */
int drgnn_net_step_max_clusters(int32_t kind, int32_t tiles, int64_t smem_bytes);
/* the same two queries for the three-layer variant (layers3 != 0) */
int64_t drgnn_net_step_smem_bytes_l(int32_t kind, int32_t tiles, int32_t F, int32_t h1, int32_t h2, int32_t max_n, int32_t max_k,
                                    int32_t max_q, int32_t max_e, int32_t Hd, int32_t out, int32_t layers3);
int drgnn_net_step_pick_tiles_l(int32_t kind, int32_t F, int32_t h1, int32_t h2, int32_t max_n, int32_t max_k, int32_t max_q,
                                int32_t max_e, int32_t Hd, int32_t out, int32_t layers3);
int drgnn_net_step(const drgnn_net_step_args* s, void* stream);
/* The same launch under the names SURVEY 8b lists for the fused per-network entry points: drgnn_sgat_step
 * requires kind == 1 (sGAT.py:62-93, 114-138), drgnn_fout_step kind == 2 (foutnet.py:56-82, 103-125) - a caller
 * that binds one network cannot launch another by a wrong `kind`.  (GINet: drgnn_ginet_step, which takes the
 * CTA-pair kernel when the graphs fit it, or drgnn_net_step with kind 0.) */
int drgnn_sgat_step(const drgnn_net_step_args* s, void* stream);
int drgnn_fout_step(const drgnn_net_step_args* s, void* stream);
/* kernels the last drgnn_net_step of this thread launched (1: reduction fused / scoring, 2: + reduction launch)
 * and the tile count it used */
int drgnn_net_step_last_launches(void);
int drgnn_net_step_last_tiles(void);
/* diagnostic: clock64 at the phase boundaries of the CTA that ran block 0 of the last launch: [0] start, [1] staged,
 * [2] zin1, [3] Z1, [4] P1, [5] zin2, [6] Z2, [7] P2, [8] read-out, [9] head, [10] head backward, [11] dZ2,
 * [12] dW2 / dzin2, [13] dP1, [14] dZ1, [15] dW1, [16] reduction.  Synchronises the device. */
int drgnn_debug_phase3_cycles(uint64_t* out32);
/* flag bit 3 of the last drgnn_net_step: %globaltimer (ns) of CTA c at kernel entry (out[2c]) and at the end of its
 * per-graph work (out[2c+1]); ctas <= 2048 */
int drgnn_debug_cta_times(uint64_t* out, int32_t ctas);

/* ---- multi-GPU: gradient exchange over NVLink peer memory fused with the optimiser ----
 * Replaces, on every rank, the sequence  [reduce per-graph rows] -> torch.distributed.all_reduce(flat
 * gradients | loss) -> torch.optim.Adam.step()  (the data-parallel form of NeuralNet.py:502-503)
 * by ONE launch: each rank stores its local sums into every rank's exchange buffer (plain P2P
 * stores), releases a per-block flag, waits for the same block of every peer, sums the `world`
 * slots in rank order (bit-identical on all ranks) and applies Adam.
 *
 * Exception to the ownership rule: the exchange region must be shareable between processes, so the
 * library allocates it (cudaMalloc, zero-filled) and exports a 64-byte CUDA IPC handle; the host
 * exchanges handles (torch.distributed.all_gather_object) and opens the peers' regions.
 * Region layout (the host computes the pointers): ctr[16] u32 (local: [0] epoch, [1] ticket,
 * [2] error status, 1 = a peer did not deliver within timeout_ns) | flags [2][world][max_blocks] u32
 * | buffers [2][world][stride] f32 | low-latency slots [2][world][stride] of 8 bytes. */
#define DRGNN_MAX_PEERS 8
#define DRGNN_IPC_HANDLE_BYTES 64
typedef struct drgnn_peer_comm {
  int32_t world; int32_t rank;
  float* xbuf[DRGNN_MAX_PEERS];        /* buffers of rank r as mapped in THIS process ([rank] = own)  */
  uint32_t* xflag[DRGNN_MAX_PEERS];    /* flags of rank r as mapped in THIS process                   */
  uint32_t* ctr;                       /* own counters                                                */
  int64_t stride;                      /* floats per slot (>= n_sum)                                  */
  int32_t max_blocks; int32_t reserved;
  uint64_t timeout_ns;                 /* watchdog of the wait (0 = 20 s)                             */
  /* low-latency slots [2][world][stride] of 8-byte words {value bits, epoch} of rank r as mapped in
   * THIS process: the in-kernel exchange of drgnn_ginet_step stores value and validity in ONE 8-byte
   * store, so it needs neither a flag nor a system-scope fence (one NVLink hop per step) */
  uint64_t* xll[DRGNN_MAX_PEERS];
} drgnn_peer_comm;
typedef struct drgnn_peer_adam_args {
  const float* partial; int32_t B; int32_t reserved0; int64_t partial_ld;  /* optional per-graph rows, summed in
                                          graph order (drgnn_ginet_step with skip_reduce); NULL: local value = grads */
  float* grads; int32_t n_params; int32_t n_sum;   /* n_sum >= n_params elements are exchanged (gradients | loss) */
  int32_t apply_adam; float lr; float beta1; float beta2; float eps; int32_t reserved1;
  float* adam_p; float* adam_m; float* adam_v; float* step_dev;
} drgnn_peer_adam_args;
int drgnn_comm_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64);
int drgnn_comm_open(const unsigned char* handle64, void** peer_ptr);
int drgnn_comm_close(void* peer_ptr);
int drgnn_comm_free(void* dev_ptr);
int drgnn_comm_status(const void* region, uint32_t* ctr4);   /* host copy of ctr[0..3] (synchronises the device) */
int drgnn_peer_reduce_adam(const drgnn_peer_comm* c, const drgnn_peer_adam_args* a, void* stream);

/* ---- multi-GPU fallback: the path's ONE collective as NCCL (SURVEY 8b "drgnn_nccl_{init,allreduce,destroy}", 8e) ----
 * For GPUs that cannot map each other's memory (no CUDA IPC / P2P: the exchange above is unavailable) the
 * data-parallel step is  [step kernel with skip_reduce = 0, fuse_adam = 0] -> drgnn_nccl_allreduce(grads | loss)
 * -> drgnn_adam_flat, the literal form of SURVEY 8e (replaces torch.distributed.all_reduce over the flat
 * gradient buffer; the reference itself is single-device, NeuralNet.py:502-503).  libdrgnn.so does not link NCCL:
 * the library is bound at run time (dlopen of $DRGNN_NCCL_LIB, "libnccl.so.2", "libnccl.so"; inside a PyTorch
 * process that is the copy torch loaded).  The 128-byte unique id of rank 0 travels to the other ranks by any
 * host channel (the Python launcher uses torch.distributed / a file); `comm` is an opaque ncclComm_t owned by
 * the caller between init and destroy; the all-reduce is in place, fp32, sum, stream-ordered. */
#define DRGNN_NCCL_ID_BYTES 128
int drgnn_nccl_available(void);                                         /* 1: an NCCL library could be bound */
int drgnn_nccl_unique_id(void* id128);                                  /* ncclGetUniqueId (rank 0)            */
int drgnn_nccl_init(void** comm, int32_t world, int32_t rank, const void* id128);   /* ncclCommInitRank (collective) */
int drgnn_nccl_allreduce(void* comm, float* buf, int64_t count, void* stream);      /* ncclAllReduce(sum, fp32), in place */
int drgnn_nccl_destroy(void* comm);                                     /* ncclCommDestroy                     */

/* ---- end-to-end feeder: the pipelined epoch loop issued from C ----
 * Replaces the per-step Python of Engine.train_batches (NeuralNet.py:490-523 over a DataLoader): for
 * step i, on the copy stream ONE host->device copy of the packed batch into staging slot `slot`, on a
 * structure stream the launch of the slot's structure-pass graph, on the main stream the launch of the
 * slot's step graph, and ONE device->host copy of [loss | predictions] (staged through a small device
 * ring on a read-back stream); a slot is rewritten only after the step that read it.  Handles: cudaStream_t / cudaGraphExec_t of the caller (PyTorch's are
 * driver handles and valid here).  No host synchronisation. */
typedef struct drgnn_feed_step {
  const void* h_src; void* d_dst; int64_t nbytes;      /* pinned packed batch -> device staging slot        */
  void* prep_graph; void* step_graph;                  /* cudaGraphExec_t of the slot's two captured graphs */
  const void* d_out; void* h_out; int64_t out_bytes;   /* [loss | pad | predictions] read-back (may be 0)   */
  int32_t slot; int32_t reserved;
} drgnn_feed_step;
/* ring (device, ring_slots x ring_stride bytes, ring_slots <= 8; 0 = read back on the main stream): the
 * step's output block is copied device-to-device into ring slot i % ring_slots on the main stream and
 * read back from there on read_stream, so the D2H latency stays off the step chain */
int drgnn_feed_run(const drgnn_feed_step* steps, int32_t n, int32_t n_slots, void* main_stream, void* copy_stream,
                   void* prep_stream0, void* prep_stream1, void* read_stream, void* ring, int64_t ring_stride,
                   int32_t ring_slots);

/* ---- GPU pre-clustering: Markov clustering of every graph of a batch, one CTA per graph ----
 * Replaces community_detection(edge_index, num_nodes, method='mcl') (community_pooling.py:95-158: networkx ->
 * scipy -> markov_clustering.run_mcl with default parameters + get_clusters), which PreCluster runs twice per
 * graph on the CPU (DataSet.py:45-88).  edge_index [2,E] int64 (idx32 = 0) or int32 with GLOBAL node ids of a
 * block-diagonal batch (graph g owns nodes node_ptr[g]..node_ptr[g+1] and edges edge_ptr[g]..edge_ptr[g+1]);
 * unit weights, undirected.  cluster [N] int64 receives per-graph LOCAL labels exactly as the reference returns
 * them (clusters sorted as member tuples, a node of several clusters keeps the last; ids may have gaps).
 * work: drgnn_mcl_work_doubles(B, max_n) doubles (two dense n x n float64 matrices per graph);
 * iters [B] (optional) receives the iterations run; status as in the structure pass. */
int64_t drgnn_mcl_work_doubles(int32_t B, int32_t max_n);
int drgnn_mcl_cluster(const int32_t* node_ptr, const int32_t* edge_ptr, const void* edge_index, int32_t B, int64_t E,
                      int32_t idx32, int32_t max_n, double* work, int64_t* cluster, int32_t* iters, int32_t* status,
                      void* stream);

/* small utilities used by the host layer */
int drgnn_relu_mask(const float* g, int32_t ldg, const float* out, int32_t ldo, int32_t rows,
                    const int32_t* rows_dev, int32_t C, float* gz, int32_t ldgz, void* stream);
int drgnn_fill_f32(float* p, float v, int64_t n, void* stream);
int drgnn_fill_i32(int32_t* p, int32_t v, int64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DRGNN_H */
