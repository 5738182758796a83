"""Tensor-level wrappers of the C-ABI (``include/drgnn.h``): PyTorch tensors in, PyTorch
tensors out, every call enqueued on the current CUDA stream, no host synchronisation
unless a function says so.  Matrices may be column slices of wider buffers (row stride =
leading dimension, unit column stride).
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (AggregateArgs, DrgnnError, GinetFusedArgs, GinetStepArgs, HeadArgs, LinearArgs, LinearWgradArgs, NetStepArgs,
                   StructureIO, call, ptr, require_cuda, stream_ptr)

I16, I32, I64, F32 = torch.int16, torch.int32, torch.int64, torch.float32


def _ld(t):
    if t.dim() == 1:
        return t.numel() if t.numel() else 1
    if t.dim() != 2 or (t.size(1) > 1 and t.stride(1) != 1):
        raise DrgnnError('expected a row-major 2-D tensor with unit column stride, got shape %s strides %s'
                         % (tuple(t.shape), tuple(t.stride())))
    return t.stride(0) if t.size(0) > 1 else max(t.stride(0), t.size(1))


def _f32(t, name):
    if t is not None and t.dtype != F32:
        raise DrgnnError('%s must be float32 (got %s)' % (name, t.dtype))
    return t


def _i32(t, name):
    if t is not None and (t.dtype != I32 or not t.is_contiguous()):
        raise DrgnnError('%s must be a contiguous int32 tensor' % name)
    return t


# ------------------------------------------------------------------------------------------
# Structure pass
# ------------------------------------------------------------------------------------------
_INT_FIELDS = ('rowptr0', 'col0', 'eid0', 'cscptr0', 'cscrow0', 'csceid0', 'cl0', 'cmptr0', 'cmem0', 'kptr0',
               'batch1', 'rowptr1', 'col1', 'cscptr1', 'cscrow1', 'csceid1', 'cl1', 'cmptr1', 'cmem1', 'kptr1',
               'batch2', 'counts', 'status', 'gstat', 'scratch_n', 'scratch_e', 'blob')
_FLOAT_FIELDS = ('w0csr', 'w0csc', 'edge_attr1', 'w1csc', 'scratch_f', 'wblob')
_I64_FIELDS = ('cl0_i64', 'batch1_i64', 'edge_index1', 'batch2_i64')


def _structure_sizes(B, N, E, L1, ne):
    n1, l1 = N + 1, L1 + 1
    ints = dict(rowptr0=n1, col0=E, eid0=E, cscptr0=n1, cscrow0=E, csceid0=E, cl0=N, cmptr0=n1, cmem0=N,
                kptr0=B + 1, batch1=N, rowptr1=n1, col1=E, cscptr1=n1, cscrow1=E, csceid1=E, cl1=L1, cmptr1=l1,
                cmem1=L1, kptr1=B + 1, batch2=L1, counts=4, status=1, gstat=8 * B,
                scratch_n=5 * (N + B + 1) + 8, scratch_e=4 * E + 8,
                blob=48 * B + 12 * N + 4 * E + 16)        # DRGNN_BLOB_WORDS: per-graph structure blobs
    floats = dict(w0csr=E if ne else 0, w0csc=E if ne else 0, edge_attr1=E * ne, w1csc=E if ne else 0,
                  scratch_f=E * max(ne, 1) if ne else 0,
                  wblob=(48 * B + 12 * N + 4 * E + 16) if ne else 0)   # edge weights parallel to the blob (sGAT)
    i64 = dict(cl0_i64=N, batch1_i64=N, edge_index1=2 * E, batch2_i64=L1)
    return ints, floats, i64


class Structure(object):
    """Everything integer that depends only on the batch: CSR/CSC of the
    level-0 graph, relabelled clusters, cluster->members CSR, pooled graph CSR/CSC, pooled
    batch vectors, level-1 clustering and the device-side counts ``[K0, E1, K1]``.
    Built by two kernel launches (``drgnn_structure_build``); shared by forward, backward
    and both GINet branches.

    The constructor sizes are CAPACITIES: one object serves every batch with
    ``B <= cap_B, N <= cap_N, E <= cap_E, L1 <= cap_L1`` at fixed device addresses (so
    captured CUDA graphs stay valid); ``B, N, E, L1`` hold the sizes of the current batch."""

    def __init__(self, B, N, E, L1, ne, device, mirrors=False):
        self.cap = (int(B), int(N), int(E), int(L1))
        self.B, self.N, self.E, self.L1, self.ne = int(B), int(N), int(E), int(L1), int(ne)
        self.device = device
        self.mirrors = mirrors
        ints, floats, i64 = _structure_sizes(self.B, self.N, self.E, self.L1, self.ne)
        pad = lambda n: (n + 3) & ~3
        tot_i = sum(pad(v) for v in ints.values())
        tot_f = sum(pad(v) for v in floats.values())
        tot_l = sum(pad(v) for v in i64.values()) if mirrors else 0
        self._iarena = torch.zeros(max(tot_i, 4), dtype=I32, device=device)   # status / counts start at zero
        self.blob_only = False
        self._farena = torch.empty(max(tot_f, 4), dtype=F32, device=device)
        self._larena = torch.empty(max(tot_l, 4), dtype=I64, device=device) if mirrors else None
        o = 0
        for k, v in ints.items():
            setattr(self, k, self._iarena[o:o + v])
            o += pad(v)
        o = 0
        for k, v in floats.items():
            setattr(self, k, self._farena[o:o + v] if v else None)
            o += pad(v)
        o = 0
        for k, v in i64.items():
            setattr(self, k, self._larena[o:o + v] if mirrors else None)
            o += pad(v) if mirrors else 0
        self._edge_attr1_flat, self._edge_index1_flat = self.edge_attr1, self.edge_index1
        self.io = StructureIO()
        for k in _INT_FIELDS + _FLOAT_FIELDS + _I64_FIELDS:
            setattr(self.io, k, ptr(getattr(self, k)))
        self._counts_host = None
        self._keep = None
        self.sticky_status = False       # True: structure_build does not zero `status` (Engine: one check per epoch)
        self._shape_views()

    def fits(self, B, N, E, L1, ne, mirrors):
        cb, cn, ce, cl = self.cap
        return B <= cb and N <= cn and E <= ce and L1 <= cl and ne == self.ne and mirrors == self.mirrors

    def _shape_views(self):
        """edge_attr1 ``[E, ne]`` and the int64 edge_index mirror ``[2, E]`` are laid out with
        the CURRENT batch's E (the kernel strides by io.E)."""
        if self.ne:
            self.edge_attr1 = self._edge_attr1_flat[:self.E * self.ne].view(self.E, self.ne)
        if self.mirrors:
            self.edge_index1 = self._edge_index1_flat[:2 * self.E].view(2, self.E)

    # device-side live counts (no sync)
    @property
    def K0_dev(self):
        return self.counts[0:1]

    @property
    def E1_dev(self):
        return self.counts[1:2]

    @property
    def K1_dev(self):
        return self.counts[2:3]

    def sync_counts(self):
        """Host copy of (K0, E1, K1) and validation of the status word.  ONE host sync."""
        if self._counts_host is None:
            host = torch.cat([self.counts, self.status]).cpu().tolist()
            st = host[4]
            if self.blob_only:
                host[:3] = [-1, -1, -1]      # the blob-only pass computes no batch totals
            if st and (self.blob_only or self.sticky_status):
                self.status.zero_()          # sticky status: reported once, then re-armed
            if st:
                msgs = [t for bit, t in _lib.STATUS_TEXT.items() if st & bit]
                raise DrgnnError('invalid batch structure: ' + '; '.join(msgs))
            self._counts_host = tuple(host[:3])
        return self._counts_host

    @property
    def K0(self):
        return self.sync_counts()[0]

    @property
    def E1(self):
        return self.sync_counts()[1]

    @property
    def K1(self):
        return self.sync_counts()[2]


def structure_build(node_ptr, edge_ptr, edge_index, cluster0, max_n, max_e, c1_ptr=None, cluster1=None,
                    edge_attr=None, clusters_are_local=True, mirrors=False, out=None, L1=None, edge_half=False):
    """Run the structure pass for one mini-batch.  ``node_ptr/edge_ptr/c1_ptr`` are int32
    ``[B+1]`` device tensors; ``edge_index`` ``[2,E]`` and ``cluster0/1`` are int64 (reference
    layout) or int32.  Returns a ``Structure`` (``out`` is reused if given)."""
    require_cuda(node_ptr, edge_ptr, edge_index, cluster0, c1_ptr, cluster1, edge_attr)
    _i32(node_ptr, 'node_ptr'), _i32(edge_ptr, 'edge_ptr'), _i32(c1_ptr, 'c1_ptr')
    B = node_ptr.numel() - 1
    N = cluster0.numel()
    E = edge_index.size(1) if edge_index.dim() == 2 else 0
    # L1: live length of cluster1 when the tensor is a capacity-sized view (packed staging buffers)
    L1 = (0 if cluster1 is None else cluster1.numel()) if L1 is None else int(L1)
    edge16 = edge_index.dtype == I16     # compact feeder batches: uint16 graph-local node ids, int32 / uint16 cluster ids
    if edge_half:                        # ... of which only the first (undirected) half of every graph travels
        if not edge16:
            raise DrgnnError('edge_half needs int16 graph-local edge ids')
        E *= 2
    if (edge16 and (cluster0.dtype not in (I32, I16) or (cluster1 is not None and cluster1.dtype != cluster0.dtype))) or \
            (not edge16 and (edge_index.dtype not in (I32, I64) or cluster0.dtype != edge_index.dtype or
                             (cluster1 is not None and cluster1.dtype != edge_index.dtype))):
        raise DrgnnError('edge_index / cluster0 / cluster1 must share one integer dtype (int64 or int32), or be '
                         'int16 graph-local edge ids with int32 cluster ids')
    if not edge_index.is_contiguous() or not cluster0.is_contiguous() or \
            (cluster1 is not None and not cluster1.is_contiguous()):
        raise DrgnnError('edge_index / cluster tensors must be contiguous')
    if L1 > N:
        raise DrgnnError('len(cluster1)=%d exceeds the number of nodes %d' % (L1, N))
    if cluster1 is not None and c1_ptr is None:
        raise DrgnnError('cluster1 given without c1_ptr')
    ne = 0
    if edge_attr is not None:
        _f32(edge_attr, 'edge_attr')
        if edge_attr.dim() == 1:
            edge_attr = edge_attr.unsqueeze(-1)          # ginet.py:54-55, sGAT.py:66-67
        if not edge_attr.is_contiguous():
            edge_attr = edge_attr.contiguous()
        ne = edge_attr.size(1)
        if edge_attr.size(0) != E:
            raise DrgnnError('edge_attr has %d rows for %d edges' % (edge_attr.size(0), E))
    s = out
    if s is None or not s.fits(B, N, E, L1, ne, mirrors):
        s = Structure(B, N, E, L1, ne, cluster0.device, mirrors)
    s.B, s.N, s.E, s.L1 = B, N, E, L1
    s._shape_views()
    s._counts_host = None
    s.blob_only = False
    s._keep = (node_ptr, edge_ptr, c1_ptr, edge_index, edge_attr, cluster0, cluster1)
    s.node_ptr, s.edge_ptr, s.c1_ptr = node_ptr, edge_ptr, c1_ptr
    s.max_n, s.max_e = int(max_n), int(max_e)
    io = s.io
    io.B, io.N, io.E, io.L1, io.ne = B, N, E, L1, ne
    io.max_n, io.max_e = int(max_n), int(max_e)
    io.clusters_are_local = 1 if clusters_are_local else 0
    io.idx32 = 2 if cluster0.dtype == I16 else (1 if cluster0.dtype == I32 else 0)
    io.edge16 = (2 if edge_half else 1) if edge16 else 0
    io.node_ptr, io.edge_ptr, io.c1_ptr = ptr(node_ptr), ptr(edge_ptr), ptr(c1_ptr)
    io.edge_index, io.edge_attr = ptr(edge_index), ptr(edge_attr)
    io.cluster0, io.cluster1 = ptr(cluster0), ptr(cluster1)
    if B == 0:
        s.counts.zero_()
        s.status.zero_()
        return s
    st = stream_ptr()
    if not s.sticky_status:
        call('drgnn_fill_i32', ptr(s.status), 0, 1, st)
    call('drgnn_structure_build', C.byref(io), st)
    return s


def structure_blob_smem(max_n, max_e, max_k=0, max_q=0, weights=True, x_words=0):
    """Shared memory (bytes) of one CTA of the blob-only structure pass for graphs of up to ``max_n`` nodes / ``max_e``
    directed edges / ``max_k``, ``max_q`` clusters of the two levels (0: ``max_n``); negative when it does not fit."""
    v = int(_lib.load().drgnn_structure_blob_smem_bytes_ex(int(max_n), int(max_e), int(max_k or 0), int(max_q or 0),
                                                           1 if weights else 0, int(x_words)))
    if v < 0 and x_words:      # the feature tile of the first aggregation is staged only when it fits
        v = int(_lib.load().drgnn_structure_blob_smem_bytes_ex(int(max_n), int(max_e), int(max_k or 0), int(max_q or 0),
                                                               1 if weights else 0, 0))
    return v


def structure_blob_fits(max_n, max_e, max_k=0, max_q=0, weights=True):
    """True when such graphs fit the blob-only structure pass (``drgnn_structure_blob``)."""
    return structure_blob_smem(max_n, max_e, max_k, max_q, weights) >= 0


ZIN_KIND = {'ginet': 0, 'sgat': 1, 'fout': 2}


def zin1_ld(kind, F):
    """Row stride of the precomputed conv1 input rows (``Structure.zin1``): the stride the step kernels keep them
    at in shared memory (F + 4 for GINet, 2F + 4 columns + 4 for the others)."""
    return (F if ZIN_KIND[kind] == 0 else 2 * F) + 4


def structure_blob(node_ptr, edge_ptr, edge_index, cluster0, max_n, max_e, c1_ptr, cluster1, out=None, L1=None,
                   edge_attr=None, x=None, zin_kind=None, edge_half=False, max_k=0, max_q=0):
    """Blob-only structure pass (``drgnn_structure_blob``): ONE launch that writes the per-graph
    structure blobs the cluster step kernel stages (graph-local indices) and nothing else - no
    global CSR arrays, no cross-graph finalize launch, no status-zeroing launch (``status`` is
    sticky: zeroed at allocation and by ``Structure.sync_counts``).  Same argument checks as
    ``structure_build``; both cluster levels are required.  ``x`` [N, F] + ``zin_kind`` ('ginet' | 'sgat' | 'fout'):
    the pass also computes the input rows of conv1's transform (the first aggregation depends on the batch only) into
    ``Structure.zin1`` for the step kernels (``drgnn_structure_io.zin1``)."""
    require_cuda(node_ptr, edge_ptr, edge_index, cluster0, c1_ptr, cluster1)
    _i32(node_ptr, 'node_ptr'), _i32(edge_ptr, 'edge_ptr'), _i32(c1_ptr, 'c1_ptr')
    if cluster1 is None or c1_ptr is None:
        raise DrgnnError('the blob-only structure pass needs cluster1 and c1_ptr')
    B = node_ptr.numel() - 1
    N = cluster0.numel()
    E = edge_index.size(1) if edge_index.dim() == 2 else 0
    L1 = cluster1.numel() if L1 is None else int(L1)
    edge16 = edge_index.dtype == I16
    if edge_half:                        # compact records: only the first (undirected) half of every graph's edges
        if not edge16:
            raise DrgnnError('edge_half needs int16 graph-local edge ids')
        E *= 2
    if (edge16 and (cluster0.dtype not in (I32, I16) or cluster1.dtype != cluster0.dtype)) or \
            (not edge16 and (edge_index.dtype not in (I32, I64) or cluster0.dtype != edge_index.dtype or
                             cluster1.dtype != edge_index.dtype)):
        raise DrgnnError('edge_index / cluster0 / cluster1 must share one integer dtype (int64 or int32), or be '
                         'int16 graph-local edge ids with int32 cluster ids')
    if not edge_index.is_contiguous() or not cluster0.is_contiguous() or not cluster1.is_contiguous():
        raise DrgnnError('edge_index / cluster tensors must be contiguous')
    if L1 > N:
        raise DrgnnError('len(cluster1)=%d exceeds the number of nodes %d' % (L1, N))
    ne = 0
    if edge_attr is not None:       # sGAT: the pass also writes the edge weights of the blob's lists (Structure.wblob)
        require_cuda(edge_attr)
        _f32(edge_attr, 'edge_attr')
        if edge_attr.dim() == 1:
            edge_attr = edge_attr.unsqueeze(-1)
        if not edge_attr.is_contiguous():
            edge_attr = edge_attr.contiguous()
        ne = edge_attr.size(1)
        if edge_attr.size(0) != E:
            raise DrgnnError('edge_attr has %d rows for %d edges' % (edge_attr.size(0), E))
    s = out
    if s is None or not s.fits(B, N, E, L1, ne if ne else s.ne, s.mirrors):
        s = Structure(B, N, E, L1, ne, cluster0.device, False)
    s.B, s.N, s.E, s.L1 = B, N, E, L1
    s._shape_views()
    s._counts_host = None
    s.blob_only = True
    s._keep = (node_ptr, edge_ptr, c1_ptr, edge_index, edge_attr, cluster0, cluster1)
    s.node_ptr, s.edge_ptr, s.c1_ptr = node_ptr, edge_ptr, c1_ptr
    s.max_n, s.max_e = int(max_n), int(max_e)
    io = s.io
    io.B, io.N, io.E, io.L1, io.ne = B, N, E, L1, ne
    io.max_n, io.max_e = int(max_n), int(max_e)
    io.clusters_are_local = 1
    io.idx32 = 2 if cluster0.dtype == I16 else (1 if cluster0.dtype == I32 else 0)
    io.edge16 = (2 if edge_half else 1) if edge16 else 0
    io.node_ptr, io.edge_ptr, io.c1_ptr = ptr(node_ptr), ptr(edge_ptr), ptr(c1_ptr)
    io.edge_index, io.edge_attr = ptr(edge_index), ptr(edge_attr)
    io.cluster0, io.cluster1 = ptr(cluster0), ptr(cluster1)
    if x is not None and zin_kind is not None:
        require_cuda(x)
        _f32(x, 'x')
        F = x.size(1)
        ld = zin1_ld(zin_kind, F)
        need = s.cap[1] * ld
        if getattr(s, '_zin1', None) is None or s._zin1.numel() < need:
            s._zin1 = torch.empty(need, dtype=F32, device=cluster0.device)     # capacity-sized: fixed address
        s.zin1, s.zin1_ld = s._zin1, ld
        s._keep = s._keep + (x,)
        io.x, io.zin1, io.F, io.ld_zin1, io.zin_kind = ptr(x), ptr(s._zin1), F, ld, ZIN_KIND[zin_kind]
    else:
        s.zin1, s.zin1_ld = None, 0
        io.x, io.zin1 = None, None
    io.max_k, io.max_q = int(max_k or 0), int(max_q or 0)   # per-graph cluster bounds: sizes of the pooled-graph bitmaps
    if B:
        call('drgnn_structure_blob', C.byref(io), stream_ptr())
    return s


def cluster_offset_(cluster, seg_ptr):
    """In-place ``get_preloaded_cluster`` (community_pooling.py:25-30) on an int64 vector."""
    require_cuda(cluster, seg_ptr)
    if cluster.dtype != I64 or not cluster.is_contiguous():
        raise DrgnnError('cluster must be a contiguous int64 tensor')
    _i32(seg_ptr, 'seg_ptr')
    B = seg_ptr.numel() - 1
    work = torch.empty(max(B, 1), dtype=I64, device=cluster.device)
    call('drgnn_cluster_offset', ptr(cluster), ptr(seg_ptr), B, ptr(work), stream_ptr())
    return cluster


def ptr_from_sorted_ids(ids, B):
    """int32 ``[B+1]`` segment pointers of a sorted int64 id vector (``batch``).  Raises (one
    host sync) if ``ids`` is not sorted or out of range."""
    require_cuda(ids)
    if ids.dtype != I64 or not ids.is_contiguous():
        raise DrgnnError('ids must be a contiguous int64 tensor')
    out = torch.empty(B + 1, dtype=I32, device=ids.device)
    status = torch.zeros(1, dtype=I32, device=ids.device)
    call('drgnn_ptr_from_sorted_ids', ptr(ids), ids.numel(), B, ptr(out), ptr(status), stream_ptr())
    if int(status.item()) != 0:
        raise DrgnnError('`batch` must be sorted ascending with values in [0, %d)' % B)
    return out


# ------------------------------------------------------------------------------------------
# Aggregation
# ------------------------------------------------------------------------------------------
def aggregate(src, rowptr, col, out, C_=None, ew=None, sscale=None, self_src=None, self_out=None, selfc_in=None,
              selfc_out=None, post_out=None, bias=None, n_rows=None, n_rows_dev=None, post_mode=0, self_mode=0,
              relu=False, tile_ptr=None, tile_eptr=None, max_tile_rows=0, max_tile_edges=0):
    """out[i] = act(selfc_i * self_src[i] + post_i * sum_p ew[p] * sscale[col[p]] * src[col[p]] + bias)
    over CSR rows (see ``drgnn_aggregate`` in include/drgnn.h).  With ``tile_ptr`` / ``tile_eptr``
    (row and CSR-slot range of every graph) the shared-memory staged per-graph kernel
    (``drgnn_aggregate_tiled``) is used."""
    require_cuda(src, rowptr, col, out, ew, sscale, self_src, self_out, selfc_in, selfc_out, post_out, bias)
    _f32(src, 'src'), _f32(out, 'out'), _i32(rowptr, 'rowptr'), _i32(col, 'col')
    a = AggregateArgs()
    a.src, a.ld_src = ptr(src), _ld(src)
    a.out, a.ld_out = ptr(out), _ld(out)
    a.rowptr, a.col = ptr(rowptr), ptr(col)
    a.ew, a.sscale = ptr(_f32(ew, 'ew')), ptr(_f32(sscale, 'sscale'))
    a.self_src, a.ld_self = ptr(_f32(self_src, 'self_src')), (_ld(self_src) if self_src is not None else 0)
    a.self_out, a.ld_self_out = ptr(_f32(self_out, 'self_out')), (_ld(self_out) if self_out is not None else 0)
    a.selfc_in, a.selfc_out = ptr(_f32(selfc_in, 'selfc_in')), ptr(_f32(selfc_out, 'selfc_out'))
    a.post_out = ptr(_f32(post_out, 'post_out'))
    a.bias = ptr(_f32(bias, 'bias'))
    a.n_rows = int(out.size(0) if n_rows is None else n_rows)
    a.n_rows_dev = ptr(n_rows_dev)
    a.C = int(out.size(1) if C_ is None else C_)
    a.post_mode, a.self_mode, a.relu = int(post_mode), int(self_mode), 1 if relu else 0
    if tile_ptr is not None:
        call('drgnn_aggregate_tiled', C.byref(a), ptr(_i32(tile_ptr, 'tile_ptr')), ptr(_i32(tile_eptr, 'tile_eptr')),
             tile_ptr.numel() - 1, int(max_tile_rows), int(max_tile_edges), stream_ptr())
    else:
        call('drgnn_aggregate', C.byref(a), stream_ptr())
    return out


# ------------------------------------------------------------------------------------------
# Dense transform
# ------------------------------------------------------------------------------------------
MATH_FMA, MATH_TF32X3, MATH_TCGEN05 = 0, 1, 2     # 2: tcgen05.mma kind::tf32 (3xTF32, TMEM accumulator), falls back to 1
default_math = MATH_FMA


def linear(X, W, Fin, Fout, out, bias=None, groups=1, w_layout=0, relu=False, out_mask=None, mask_scale=1.0,
           rows=None, rows_dev=None, math=None):
    """Y[r, g*Fout+o] = act(sum_k X[r, g*Fin+k] * W_g[k,o] + bias) (* mask) - ``drgnn_linear``."""
    require_cuda(X, W, out, bias, out_mask)
    _f32(X, 'X'), _f32(W, 'W'), _f32(out, 'out'), _f32(bias, 'bias'), _f32(out_mask, 'out_mask')
    if not W.is_contiguous() or W.numel() != groups * Fin * Fout:
        raise DrgnnError('W must be contiguous with groups*Fin*Fout = %d elements (got %d)'
                         % (groups * Fin * Fout, W.numel()))
    a = LinearArgs()
    a.X, a.ldx = ptr(X), _ld(X)
    a.W, a.bias = ptr(W), ptr(bias)
    a.Y, a.ldy = ptr(out), _ld(out)
    a.out_mask, a.ld_mask = ptr(out_mask), (_ld(out_mask) if out_mask is not None else 0)
    a.mask_scale = float(mask_scale)
    a.rows = int(out.size(0) if rows is None else rows)
    a.rows_dev = ptr(rows_dev)
    a.Fin, a.Fout, a.groups = int(Fin), int(Fout), int(groups)
    a.w_layout, a.relu = int(w_layout), 1 if relu else 0
    a.math = int(default_math if math is None else math)
    call('drgnn_linear', C.byref(a), stream_ptr())
    return out


def linear_wgrad_work_floats(rows, Fin, Fout, groups=1):
    return int(_lib.load().drgnn_linear_wgrad_work_floats(int(rows), int(Fin), int(Fout), int(groups)))


def linear_wgrad(X, G, Fin, Fout, dW, dbias=None, groups=1, w_layout=0, rows=None, rows_dev=None, accumulate=False,
                 work=None):
    """dW (+)= G^T X (layout as the weight), dbias (+)= column sums of G - ``drgnn_linear_wgrad``."""
    require_cuda(X, G, dW, dbias, work)
    _f32(X, 'X'), _f32(G, 'G'), _f32(dW, 'dW'), _f32(dbias, 'dbias')
    rows = int(G.size(0) if rows is None else rows)
    need = linear_wgrad_work_floats(rows, Fin, Fout, groups)
    if work is None:
        work = torch.zeros(need, dtype=F32, device=G.device)      # zero: the fast path keeps a ticket counter in it
    if not dW.is_contiguous() or dW.numel() != groups * Fin * Fout:
        raise DrgnnError('dW must be contiguous with groups*Fin*Fout elements')
    a = LinearWgradArgs()
    a.X, a.ldx = ptr(X), _ld(X)
    a.G, a.ldg = ptr(G), _ld(G)
    a.dW, a.dbias = ptr(dW), ptr(dbias)
    a.rows, a.rows_dev = rows, ptr(rows_dev)
    a.Fin, a.Fout, a.groups = int(Fin), int(Fout), int(groups)
    a.w_layout, a.accumulate = int(w_layout), 1 if accumulate else 0
    a.work, a.work_floats = ptr(work), work.numel()
    call('drgnn_linear_wgrad', C.byref(a), stream_ptr())
    return dW


# ------------------------------------------------------------------------------------------
# Pooling / read-out
# ------------------------------------------------------------------------------------------
def maxpool_fwd(x, cmptr, cmem, out, argmax, n_clusters=None, n_clusters_dev=None):
    require_cuda(x, cmptr, cmem, out, argmax)
    _f32(x, 'x'), _f32(out, 'out'), _i32(cmptr, 'cmptr'), _i32(cmem, 'cmem'), _i32(argmax, 'argmax')
    Cc = x.size(1)
    n = int(out.size(0) if n_clusters is None else n_clusters)
    call('drgnn_maxpool_fwd', ptr(x), _ld(x), ptr(cmptr), ptr(cmem), n, ptr(n_clusters_dev), Cc, ptr(out), _ld(out),
         ptr(argmax), stream_ptr())
    return out, argmax


def maxpool_bwd(g, argmax, cl, dx, relu_out=None, n_nodes=None, n_nodes_dev=None):
    require_cuda(g, argmax, cl, dx, relu_out)
    _f32(g, 'g'), _f32(dx, 'dx'), _f32(relu_out, 'relu_out'), _i32(argmax, 'argmax'), _i32(cl, 'cl')
    Cc = g.size(1)
    n = int(dx.size(0) if n_nodes is None else n_nodes)
    call('drgnn_maxpool_bwd', ptr(g), _ld(g), ptr(argmax), ptr(cl), ptr(relu_out),
         _ld(relu_out) if relu_out is not None else 0, n, ptr(n_nodes_dev), Cc, ptr(dx), _ld(dx), stream_ptr())
    return dx


def segment_mean_fwd(x, seg_ptr, out):
    require_cuda(x, seg_ptr, out)
    _f32(x, 'x'), _f32(out, 'out'), _i32(seg_ptr, 'seg_ptr')
    call('drgnn_segment_mean_fwd', ptr(x), _ld(x), ptr(seg_ptr), seg_ptr.numel() - 1, x.size(1), ptr(out), _ld(out),
         stream_ptr())
    return out


def segment_mean_bwd(g, seg_ptr, dx):
    require_cuda(g, seg_ptr, dx)
    _f32(g, 'g'), _f32(dx, 'dx'), _i32(seg_ptr, 'seg_ptr')
    call('drgnn_segment_mean_bwd', ptr(g), _ld(g), ptr(seg_ptr), seg_ptr.numel() - 1, g.size(1), ptr(dx), _ld(dx),
         stream_ptr())
    return dx


# ------------------------------------------------------------------------------------------
# Loss / optimiser / utilities
# ------------------------------------------------------------------------------------------
def mse_loss(pred, y, inv_B_global, loss_out, dpred=None, sigmoid=False):
    require_cuda(pred, y, loss_out, dpred)
    call('drgnn_mse_loss', ptr(_f32(pred, 'pred')), ptr(_f32(y, 'y')), pred.numel(), float(inv_B_global),
         1 if sigmoid else 0, ptr(loss_out), ptr(dpred), stream_ptr())
    return loss_out


def ce_loss(logits, target, inv_norm_global, loss_out, dlogits=None, class_w=None):
    require_cuda(logits, target, loss_out, dlogits, class_w)
    if target.dtype != I64:
        raise DrgnnError('target must be int64 class indices')
    call('drgnn_ce_loss', ptr(_f32(logits, 'logits')), _ld(logits), ptr(target), ptr(_f32(class_w, 'class_w')),
         logits.size(0), logits.size(1), float(inv_norm_global), ptr(loss_out), ptr(dlogits), stream_ptr())
    return loss_out


def adam_flat(param, grad, exp_avg, exp_avg_sq, step_dev, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_scale=1.0):
    require_cuda(param, grad, exp_avg, exp_avg_sq, step_dev)
    for t in (param, grad, exp_avg, exp_avg_sq):
        if t.dtype != F32 or not t.is_contiguous():
            raise DrgnnError('adam_flat works on contiguous float32 buffers')
    call('drgnn_adam_flat', ptr(param), ptr(grad), ptr(exp_avg), ptr(exp_avg_sq), ptr(step_dev), param.numel(),
         float(lr), float(beta1), float(beta2), float(eps), float(grad_scale), stream_ptr())


def ginet_fused_fits(F, h1, h2, nb, max_n, max_k, max_q):
    """True if one graph of these bounds fits the shared memory of the fused forward AND backward."""
    f = _lib.load().drgnn_ginet_fused_smem_bytes
    a = [int(v) for v in (F, h1, h2, nb, max_n, max_k, max_q)]
    return int(f(*a, 0)) >= 0 and int(f(*a, 1)) >= 0


def ginet_fused_args(st, x, W1, W2, Zin1, Z1, arg0, Zin2, Z2, arg1, R, node_ptr, B, F, h1, h2, nb, max_n, max_k, max_q,
                     dR=None, partial=None, dW1=None, dW2=None):
    """Build the argument block of ``drgnn_ginet_fused_{fwd,bwd}`` from a ``Structure``."""
    require_cuda(x, W1, W2, Zin1, Z1, arg0, Zin2, Z2, arg1, R, node_ptr, dR, partial, dW1, dW2)
    for t_, n_ in ((x, 'x'), (Zin1, 'Zin1'), (Z1, 'Z1'), (Zin2, 'Zin2'), (Z2, 'Z2'), (R, 'R')):
        if not t_.is_contiguous() or t_.dtype != F32:
            raise DrgnnError('%s must be a contiguous float32 tensor' % n_)
    a = GinetFusedArgs()
    a.B, a.F, a.h1, a.h2, a.nb = int(B), int(F), int(h1), int(h2), int(nb)
    a.max_n, a.max_k, a.max_q = int(max_n), int(max_k), int(max_q)
    a.node_ptr = ptr(node_ptr)
    a.rowptr0, a.col0, a.rowptr1, a.col1 = ptr(st.rowptr0), ptr(st.col0), ptr(st.rowptr1), ptr(st.col1)
    a.cscptr1, a.cscrow1 = ptr(st.cscptr1), ptr(st.cscrow1)
    a.cmptr0, a.cmem0, a.cl0, a.kptr0 = ptr(st.cmptr0), ptr(st.cmem0), ptr(st.cl0), ptr(st.kptr0)
    a.cmptr1, a.cmem1, a.cl1, a.kptr1 = ptr(st.cmptr1), ptr(st.cmem1), ptr(st.cl1), ptr(st.kptr1)
    a.status = ptr(st.status)
    a.W1, a.W2, a.x = ptr(W1), ptr(W2), ptr(x)
    a.Zin1, a.Z1, a.arg0, a.Zin2, a.Z2, a.arg1, a.R = ptr(Zin1), ptr(Z1), ptr(arg0), ptr(Zin2), ptr(Z2), ptr(arg1), ptr(R)
    a.dR, a.partial, a.dW1, a.dW2 = ptr(dR), ptr(partial), ptr(dW1), ptr(dW2)
    return a


def ginet_fused_fwd(a):
    call('drgnn_ginet_fused_fwd', C.byref(a), stream_ptr())


def ginet_fused_bwd(a):
    call('drgnn_ginet_fused_bwd', C.byref(a), stream_ptr())


def ginet_step_fits(F, h1, h2, nb, max_n, max_k, max_q, Hd, out):
    a = [int(v) for v in (F, h1, h2, nb, max_n, max_k, max_q, Hd, out)]
    return int(_lib.load().drgnn_ginet_step_smem_bytes(*a)) >= 0


def ginet_step(fa, fc1_w, fc1_b, fc2_w, fc2_b, pred, task=0, inv_norm=1.0, y=None, y_class=None, class_w=None, keep=None,
               keep_scale=1.0, loss=None, partial=None, grads=None, n_params=0, offsets=None, forward_only=False,
               drop_p=0.0, seed=0, step_dev=None, adam=None, skip_reduce=False, max_e=0, mirror=False, variant=0,
               fuse_reduce=True, blob=None, edge_ptr=None, comm=None, gdesc=None, tc=False, timers=False, zin1=None, head2=False):
    """Whole GINet step of every graph in one launch (``drgnn_ginet_step``); ``fa`` from
    ``ginet_fused_args``.  ``max_e`` (directed edges of the largest graph) enables the cluster
    kernel (a pair of CTAs per graph, everything in shared memory); ``mirror`` makes it store the
    intermediates to global memory too; ``variant`` 1 / 2 forces the single-CTA / cluster kernel."""
    require_cuda(fc1_w, fc1_b, fc2_w, fc2_b, pred, y, y_class, class_w, keep, loss, partial, grads)
    s = GinetStepArgs()
    s.g = fa
    s.fc1_w, s.fc1_b, s.fc2_w, s.fc2_b = ptr(fc1_w), ptr(fc1_b), ptr(fc2_w), ptr(fc2_b)
    s.Hd, s.out = fc1_w.size(0), fc2_w.size(0)
    s.keep, s.keep_scale = ptr(keep), float(keep_scale) if keep is not None else 1.0
    s.y, s.y_class, s.class_w = ptr(y), ptr(y_class), ptr(class_w)
    s.task, s.inv_norm = int(task), float(inv_norm)
    s.pred, s.loss = ptr(pred), ptr(loss)
    s.partial, s.partial_ld = ptr(partial), (partial.stride(0) if partial is not None else 0)
    s.grads, s.n_params = ptr(grads), int(n_params)
    if offsets is not None:
        s.off_w1, s.off_w2, s.off_fc1w, s.off_fc1b, s.off_fc2w, s.off_fc2b = [int(o) for o in offsets]
    s.forward_only = 1 if forward_only else 0
    s.drop_p, s.seed, s.step_dev = float(drop_p), int(seed) & 0xffffffff, ptr(step_dev)
    if keep is None and drop_p > 0:
        s.keep_scale = float(keep_scale)
    if adam is not None:        # dict(p, m, v, lr, beta1, beta2, eps): fuse the optimiser into the reduction launch
        require_cuda(adam['p'], adam['m'], adam['v'], step_dev)
        s.fuse_adam = 1
        s.adam_p, s.adam_m, s.adam_v = ptr(adam['p']), ptr(adam['m']), ptr(adam['v'])
        s.lr, s.beta1, s.beta2, s.eps = float(adam['lr']), float(adam['beta1']), float(adam['beta2']), float(adam['eps'])
    s.skip_reduce = 1 if skip_reduce else 0
    require_cuda(blob, edge_ptr)
    s.blob, s.edge_ptr = ptr(blob), ptr(edge_ptr)
    # comm (parallel.PeerComm): the gradient exchange over peer memory runs inside the cluster kernel
    s.comm = C.addressof(comm.struct) if comm is not None else None
    require_cuda(gdesc)
    s.gdesc = ptr(gdesc)          # per-graph extents of a blob-only structure pass (Structure.gstat)
    require_cuda(zin1)
    s.zin1 = ptr(zin1)            # AX of the structure pass (Structure.zin1, row stride F + 4)
    # flags: bit 0 mirror the intermediates, bit 1 no in-kernel reduction, bit 2 dense products on mma.sync 3xTF32 tiles,
    # bit 3 phase clocks of block 0 (drgnn_debug_phase_cycles)
    s.max_e, s.variant = int(max_e or 0), int(variant)
    # bit 6: head v2 of the cluster kernel (fc2 / loss / dLoss in every warp; fc1.weight gradient rows formed by the
    # in-kernel reduction from the fc1.bias gradient and read-out rows instead of being stored; same bits)
    s.flags = (1 if mirror else 0) | (0 if fuse_reduce else 2) | (4 if tc else 0) | (8 if timers else 0) | (16 if int(timers) > 1 else 0) \
        | (64 if head2 else 0)
    call('drgnn_ginet_step', C.byref(s), stream_ptr())
    # KERNELS_PER_CALL counts 2 (per-graph kernel + reduction); scoring, the peer exchange and the
    # in-kernel reduction (grid barrier) launch only the per-graph kernel
    _lib.kernel_count += int(_lib.load().drgnn_ginet_step_last_launches()) - 2


def ginet_step_last_variant():
    """1 = single-CTA kernel, 2 = cluster kernel: what the last ``ginet_step`` of this thread launched."""
    return int(_lib.load().drgnn_ginet_step_last_variant())


def ginet_step2_max_clusters(smem_bytes):
    """Co-resident 2-CTA clusters of the cluster step kernel at ``smem_bytes`` per CTA (needs a device)."""
    return int(_lib.load().drgnn_ginet_step2_max_clusters(int(smem_bytes)))


def ginet_step2_smem_bytes(F, h1, h2, max_n, max_k, max_q, max_e, Hd, out):
    return int(_lib.load().drgnn_ginet_step2_smem_bytes(*[int(v) for v in (F, h1, h2, max_n, max_k, max_q, max_e, Hd, out)]))


def mcl_cluster(edge_index, node_ptr, edge_ptr, max_n, return_iters=False):
    """Markov clustering of every graph of a block-diagonal batch on the GPU (``drgnn_mcl_cluster``): int64
    ``[N]`` per-graph local labels as ``community_detection(..., method='mcl')`` returns them
    (community_pooling.py:142-155).  ``edge_index [2,E]`` int64 / int32 with global node ids.  ONE host sync
    (status check)."""
    require_cuda(edge_index, node_ptr, edge_ptr)
    _i32(node_ptr, 'node_ptr'), _i32(edge_ptr, 'edge_ptr')
    if edge_index.dtype not in (I32, I64) or edge_index.dim() != 2 or edge_index.size(0) != 2:
        raise DrgnnError('edge_index must be an int64 / int32 [2, E] tensor')
    edge_index = edge_index.contiguous()
    B = node_ptr.numel() - 1
    dev = edge_index.device
    N = int(node_ptr[-1].item()) if B > 0 else 0
    cluster = torch.zeros(N, dtype=I64, device=dev)
    iters = torch.zeros(max(B, 1), dtype=I32, device=dev)
    status = torch.zeros(1, dtype=I32, device=dev)
    need = int(_lib.load().drgnn_mcl_work_doubles(B, int(max_n)))
    work = torch.empty(max(need, 1), dtype=torch.float64, device=dev)
    call('drgnn_mcl_cluster', ptr(node_ptr), ptr(edge_ptr), ptr(edge_index), B, edge_index.size(1),
         1 if edge_index.dtype == I32 else 0, int(max_n), ptr(work), ptr(cluster), ptr(iters), ptr(status), stream_ptr())
    st = int(status.item())
    if st:
        raise DrgnnError('mcl_cluster: ' + '; '.join(t for bit, t in _lib.STATUS_TEXT.items() if st & bit))
    return (cluster, iters[:B]) if return_iters else cluster


NET_KINDS = {'ginet': 0, 'sgat': 1, 'fout': 2}


def net_step_pick_tiles(kind, F, h1, h2, max_n, max_k, max_q, max_e, Hd, out, layers3=False):
    """Smallest number of node tiles (CTAs sharing one graph) for which the general cluster step kernel fits
    shared memory; 0 when the graph fits no cluster of 8 CTAs."""
    v = int(_lib.load().drgnn_net_step_pick_tiles_l(*[int(x) for x in (NET_KINDS[kind], F, h1, h2, max_n, max_k, max_q, max_e,
                                                                         Hd, out, 1 if layers3 else 0)]))
    return v if v > 0 else 0


def net_step_smem_bytes(kind, tiles, F, h1, h2, max_n, max_k, max_q, max_e, Hd, out, layers3=False):
    return int(_lib.load().drgnn_net_step_smem_bytes_l(*[int(x) for x in (NET_KINDS[kind], tiles, F, h1, h2, max_n, max_k,
                                                                           max_q, max_e, Hd, out, 1 if layers3 else 0)]))


def net_step_max_clusters(kind, tiles, smem_bytes):
    """Clusters of the general step kernel the device holds at once (needs a device)."""
    return int(_lib.load().drgnn_net_step_max_clusters(NET_KINDS[kind], int(tiles), int(smem_bytes)))


def net_step(kind, st, x, params, offsets, B, F, h1, h2, Hd, out, max_n, max_e, max_k, max_q, pred, node_ptr, edge_ptr,
             tiles=0, task=0, inv_norm=1.0, y=None, y_class=None, class_w=None, keep=None, keep_scale=1.0, drop_p=0.0,
             seed=0, loss=None, R=None, partial=None, grads=None, n_params=0, forward_only=False, step_dev=None, adam=None,
             skip_reduce=False, fuse_reduce=True, comm=None, mirror=None, tc=False, timers=False, lpt=False, zin1=None):
    """Whole step of every graph in one launch for GINet / sGAT / FoutNet (``drgnn_net_step``, the general
    cluster kernel).  ``st``: the ``Structure`` of the batch (blob, + wblob for sGAT); ``offsets``: dict of the
    tensor offsets inside the flat parameter buffer (w1, b1, w2, b2, fc1w, fc1b, fc2w, fc2b; b1 / b2 None for
    GINet); ``mirror``: dict(Zin1, Z1, arg0, Zin2, Z2, arg1) of global buffers that receive the intermediates
    (tests; needs a full structure pass for kptr0 / kptr1)."""
    require_cuda(x, params, pred, y, y_class, class_w, keep, loss, R, partial, grads, step_dev, node_ptr, edge_ptr)
    s = NetStepArgs()
    s.kind = NET_KINDS[kind]
    s.B, s.F, s.h1, s.h2, s.Hd, s.out = int(B), int(F), int(h1), int(h2), int(Hd), int(out)
    s.max_n, s.max_e, s.max_k, s.max_q = int(max_n), int(max_e), int(max_k), int(max_q)
    s.tiles = int(tiles)
    s.x = ptr(_f32(x, 'x'))
    s.blob = ptr(st.blob)
    s.wblob = ptr(st.wblob) if kind == 'sgat' else None
    s.gdesc = ptr(st.gstat) if st.blob_only else None
    s.node_ptr, s.edge_ptr = ptr(_i32(node_ptr, 'node_ptr')), ptr(_i32(edge_ptr, 'edge_ptr'))
    s.params = ptr(_f32(params, 'params'))
    o = offsets
    s.off_w1, s.off_w2 = int(o['w1']), int(o['w2'])
    s.off_b1 = -1 if o.get('b1') is None else int(o['b1'])
    s.off_b2 = -1 if o.get('b2') is None else int(o['b2'])
    s.off_fc1w, s.off_fc1b, s.off_fc2w, s.off_fc2b = int(o['fc1w']), int(o['fc1b']), int(o['fc2w']), int(o['fc2b'])
    if o.get('w3') is not None:          # three-layer variant (conv3 on the coarsened graph)
        s.layers3, s.off_w3, s.off_b3 = 1, int(o['w3']), int(o['b3'])
    s.keep, s.keep_scale = ptr(_f32(keep, 'keep')), float(keep_scale)
    s.drop_p, s.seed = float(drop_p), int(seed) & 0xffffffff
    s.y, s.y_class, s.class_w = ptr(y), ptr(y_class), ptr(class_w)
    s.task, s.inv_norm = int(task), float(inv_norm)
    s.forward_only, s.skip_reduce = 1 if forward_only else 0, 1 if skip_reduce else 0
    s.pred, s.loss, s.R = ptr(pred), ptr(loss), ptr(R)
    s.partial, s.partial_ld = ptr(partial), (partial.stride(0) if partial is not None else 0)
    s.grads, s.n_params = ptr(grads), int(n_params)
    if adam is not None:
        require_cuda(adam['p'], adam['m'], adam['v'])
        s.fuse_adam = 1
        s.adam_p, s.adam_m, s.adam_v = ptr(adam['p']), ptr(adam['m']), ptr(adam['v'])
        s.lr, s.beta1, s.beta2, s.eps = float(adam['lr']), float(adam['beta1']), float(adam['beta2']), float(adam['eps'])
    s.step_dev = ptr(step_dev)
    s.status = ptr(st.status)
    s.comm = C.addressof(comm.struct) if comm is not None else None
    require_cuda(zin1)
    s.zin1 = ptr(zin1)            # conv1 input rows of the structure pass (Structure.zin1, row stride Kin1 + 4)
    # bit 2: dense products on mma.sync 3xTF32 tiles; bit 3: phase clocks of block 0 (drgnn_debug_phase3_cycles)
    # bit 5: clusters take the graphs in descending size order (grids larger than the device, mixed sizes)
    s.flags = (0 if fuse_reduce else 2) | (4 if tc else 0) | (8 if timers else 0) | (32 if lpt else 0)
    if mirror is not None:
        require_cuda(*mirror.values())
        s.flags |= 1
        s.kptr0, s.kptr1 = ptr(st.kptr0), ptr(st.kptr1)
        s.Zin1, s.Z1, s.arg0 = ptr(mirror['Zin1']), ptr(mirror['Z1']), ptr(mirror['arg0'])
        s.Zin2, s.Z2, s.arg1 = ptr(mirror['Zin2']), ptr(mirror['Z2']), ptr(mirror['arg1'])
    # the per-network names of SURVEY 8b (same launch; the entry point checks the kind it is bound for)
    call(('drgnn_net_step', 'drgnn_sgat_step', 'drgnn_fout_step')[int(s.kind)], C.byref(s), stream_ptr())
    _lib.kernel_count += int(_lib.load().drgnn_net_step_last_launches()) - 1


def net_step_last():
    """(kernels launched, node tiles) of the last ``net_step`` of this thread."""
    lib = _lib.load()
    return int(lib.drgnn_net_step_last_launches()), int(lib.drgnn_net_step_last_tiles())


def peer_reduce_adam(comm, grads, n_params, n_sum, partial=None, B=0, adam=None, step_dev=None):
    """Gradient exchange over NVLink peer memory + rank-ordered sum + Adam in ONE launch
    (``drgnn_peer_reduce_adam``).  ``comm``: ``parallel.PeerComm``; ``partial``: optional per-graph
    rows of the whole-step kernel (summed in graph order first); ``grads [>= n_sum]`` receives the
    global sums (gradients | loss)."""
    require_cuda(grads, partial, step_dev)
    a = _lib.PeerAdamArgs()
    a.partial, a.B = ptr(partial), int(B)
    a.partial_ld = partial.stride(0) if partial is not None else 0
    a.grads, a.n_params, a.n_sum = ptr(grads), int(n_params), int(n_sum)
    a.step_dev = ptr(step_dev)
    if adam is not None:
        require_cuda(adam['p'], adam['m'], adam['v'])
        a.apply_adam = 1
        a.adam_p, a.adam_m, a.adam_v = ptr(adam['p']), ptr(adam['m']), ptr(adam['v'])
        a.lr, a.beta1, a.beta2, a.eps = float(adam['lr']), float(adam['beta1']), float(adam['beta2']), float(adam['eps'])
    call('drgnn_peer_reduce_adam', C.byref(comm.struct), C.byref(a), stream_ptr())


TASK_NONE, TASK_MSE, TASK_MSE_SIGMOID, TASK_CE = 0, 1, 2, 3


def head_fits(C_, Hd, out):
    return int(_lib.load().drgnn_head_smem_bytes(int(C_), int(Hd), int(out))) >= 0


def head(R, W1, b1, W2, b2, pred, task=TASK_NONE, inv_norm=1.0, y=None, y_class=None, class_w=None, keep=None,
         keep_scale=1.0, loss=None, dW1=None, db1=None, dW2=None, db2=None, dR=None, H=None):
    """Fused fc1 -> fc2 -> loss -> backward of both on the B read-out rows (``drgnn_head``)."""
    require_cuda(R, W1, b1, W2, b2, pred, y, y_class, class_w, keep, loss, dW1, db1, dW2, db2, dR, H)
    for t_, n_ in ((R, 'R'), (W1, 'W1'), (b1, 'b1'), (W2, 'W2'), (b2, 'b2'), (pred, 'pred'), (y, 'y'), (keep, 'keep')):
        _f32(t_, n_)
    a = HeadArgs()
    a.R, a.ldr = ptr(R), _ld(R)
    a.W1, a.b1, a.W2, a.b2 = ptr(W1), ptr(b1), ptr(W2), ptr(b2)
    a.keep, a.keep_scale = ptr(keep), float(keep_scale) if keep is not None else 1.0
    a.y, a.y_class, a.class_w = ptr(y), ptr(y_class), ptr(class_w)
    a.B, a.C, a.Hd, a.out = R.size(0), W1.size(1), W1.size(0), W2.size(0)
    a.task, a.inv_norm = int(task), float(inv_norm)
    a.pred, a.loss, a.H = ptr(pred), ptr(loss), ptr(H)
    a.dW1, a.db1, a.dW2, a.db2 = ptr(dW1), ptr(db1), ptr(dW2), ptr(db2)
    a.dR, a.lddr = ptr(dR), (_ld(dR) if dR is not None else 0)
    call('drgnn_head', C.byref(a), stream_ptr())
    return pred


def relu_mask(g, out, gz, rows=None, rows_dev=None):
    require_cuda(g, out, gz)
    n = int(g.size(0) if rows is None else rows)
    call('drgnn_relu_mask', ptr(_f32(g, 'g')), _ld(g), ptr(_f32(out, 'out')), _ld(out), n, ptr(rows_dev), g.size(1),
         ptr(_f32(gz, 'gz')), _ld(gz), stream_ptr())
    return gz


def fill_(t, value):
    require_cuda(t)
    if not t.is_contiguous():
        raise DrgnnError('fill_ needs a contiguous tensor')
    if t.dtype == F32:
        call('drgnn_fill_f32', ptr(t), float(value), t.numel(), stream_ptr())
    elif t.dtype == I32:
        call('drgnn_fill_i32', ptr(t), int(value), t.numel(), stream_ptr())
    else:
        raise DrgnnError('fill_ supports float32 / int32')
    return t
