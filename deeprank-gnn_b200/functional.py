"""``torch.autograd.Function`` shims over the C-ABI kernels, used by the drop-in ``nn.Module``
classes (``ginet.GINet``, ``sGAT.sGAT``, ``foutnet.FoutNet``) and available to user-defined
networks (reference "Custom GNN" tutorial, docs/tutorial.advanced.rst).  They are the only
callers of the library besides the fused engine.

Each function is the pair (forward kernel, hand-written backward kernel) of SURVEY 8a-bis.
"""
import torch

from . import ops

F32, I32 = torch.float32, torch.int32


class GraphOp(object):
    """CSR (by destination) and CSC (by source) of one directed edge set, int32, on the device."""

    def __init__(self, rowptr, col, cscptr, cscrow, n_rows, w_csr=None, w_csc=None):
        self.rowptr, self.col, self.cscptr, self.cscrow = rowptr, col, cscptr, cscrow
        self.n_rows = int(n_rows)
        self.w_csr, self.w_csc = w_csr, w_csc

    @staticmethod
    def from_edge_index(edge_index, num_nodes, edge_weight=None):
        """For stand-alone layer calls on a bare ``edge_index`` (no graph pointers known): the
        ordering uses torch's stable sort; network forwards use the structure pass instead."""
        row, col = edge_index[0], edge_index[1]
        n = int(num_nodes)
        dev = edge_index.device

        def build(key, other):
            order = torch.argsort(key, stable=True)
            ptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            ptr[1:] = torch.cumsum(torch.bincount(key, minlength=n), 0)
            return ptr.to(I32), other[order].to(I32).contiguous(), order
        rowptr, c, o1 = build(row, col)
        cscptr, r, o2 = build(col, row)
        w1 = w2 = None
        if edge_weight is not None:
            w = edge_weight.reshape(-1).to(F32)
            w1, w2 = w[o1].contiguous(), w[o2].contiguous()
        return GraphOp(rowptr, c, cscptr, r, n, w1, w2)


class _AggSum(torch.autograd.Function):
    """Z = A X (GINet aggregation, ginet.py:57-71 with alpha == 1);  dX = A^T dZ."""

    @staticmethod
    def forward(ctx, x, gop):
        ctx.gop = gop
        ctx.n_src = x.size(0)
        out = torch.empty(gop.n_rows, x.size(1), dtype=F32, device=x.device)
        ops.aggregate(x.contiguous(), gop.rowptr, gop.col, out)
        return out

    @staticmethod
    def backward(ctx, g):
        gop = ctx.gop
        dx = torch.empty(ctx.n_src, g.size(1), dtype=F32, device=g.device)
        ops.aggregate(g.contiguous(), gop.cscptr, gop.cscrow, dx)
        return dx, None


class _MeanConcat(torch.autograd.Function):
    """[self_i * x_i || mean_i]: the input of the sGAT / Fout transforms.
    weighted (sGAT.py:62-93): mean_i = sum_e a_e x_col / max(deg,1), self_i = sum_e a_e / max(deg,1)
    plain (foutnet.py:56-82): mean_i = sum_e x_col / deg (NaN for deg 0), self_i = 1."""

    @staticmethod
    def forward(ctx, x, gop, weighted):
        x = x.contiguous()
        n, C = gop.n_rows, x.size(1)
        out = torch.empty(n, 2 * C, dtype=F32, device=x.device)
        s = torch.empty(n, dtype=F32, device=x.device)
        post = torch.empty(n, dtype=F32, device=x.device)
        if weighted:
            ops.aggregate(x, gop.rowptr, gop.col, out[:, C:], C_=C, ew=gop.w_csr, self_src=x, self_out=out[:, :C],
                          selfc_out=s, post_out=post, post_mode=1, self_mode=2)
        else:
            ops.aggregate(x, gop.rowptr, gop.col, out[:, C:], C_=C, self_src=x, self_out=out[:, :C], post_out=post,
                          post_mode=2, self_mode=1)
        ctx.gop, ctx.weighted, ctx.C = gop, weighted, C
        ctx.save_for_backward(s, post)
        return out

    @staticmethod
    def backward(ctx, g):
        s, post = ctx.saved_tensors
        gop, C = ctx.gop, ctx.C
        g = g.contiguous()
        dx = torch.empty(g.size(0), C, dtype=F32, device=g.device)
        if ctx.weighted:
            ops.aggregate(g[:, C:], gop.cscptr, gop.cscrow, dx, C_=C, ew=gop.w_csc, sscale=post, self_src=g[:, :C],
                          selfc_in=s, self_mode=3)
        else:
            ops.aggregate(g[:, C:], gop.cscptr, gop.cscrow, dx, C_=C, sscale=post, self_src=g[:, :C], self_mode=1)
        return dx, None, None


class _Linear(torch.autograd.Function):
    """Y = act(X W + b) on ``groups`` column blocks; backward = masked dY -> (dW, db) + dX."""

    @staticmethod
    def forward(ctx, X, W, bias, Fin, Fout, groups, w_layout, relu, keep_mask, keep_scale):
        X = X.contiguous()
        Wc = W.contiguous()
        out = torch.empty(X.size(0), groups * Fout, dtype=F32, device=X.device)
        ops.linear(X, Wc.view(-1), Fin, Fout, out, bias=bias, groups=groups, w_layout=w_layout, relu=relu,
                   out_mask=keep_mask, mask_scale=keep_scale if keep_mask is not None else 1.0)
        ctx.cfg = (Fin, Fout, groups, w_layout, relu, keep_scale if keep_mask is not None else 1.0)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(X, Wc, out if (relu or keep_mask is not None) else None)
        return out

    @staticmethod
    def backward(ctx, g):
        X, W, out = ctx.saved_tensors
        Fin, Fout, groups, w_layout, relu, scale = ctx.cfg
        g = g.contiguous()
        if out is not None:
            # fused ReLU / dropout gate: dY * (Y > 0) * scale
            gz = torch.empty_like(g)
            ops.relu_mask(g, out, gz)
            if scale != 1.0:
                gz = gz * scale
            g = gz
        dW = torch.empty(W.numel(), dtype=F32, device=g.device)
        db = torch.empty(groups * Fout, dtype=F32, device=g.device) if ctx.has_bias else None
        ops.linear_wgrad(X, g, Fin, Fout, dW, db, groups=groups, w_layout=w_layout)
        dX = None
        if ctx.needs_input_grad[0]:
            dX = torch.empty(X.size(0), groups * Fin, dtype=F32, device=g.device)
            ops.linear(g, W.view(-1), Fout, Fin, dX, groups=groups, w_layout=1 - w_layout)
        return dX, dW.view(W.shape), db, None, None, None, None, None, None, None


class _MaxPool(torch.autograd.Function):
    """torch_scatter.scatter_max over clusters (community_pooling.py:201, max_pool_x): first
    member wins ties, gradient routed to the argmax only."""

    @staticmethod
    def forward(ctx, x, cmptr, cmem, cl, K):
        x = x.contiguous()
        out = torch.empty(K, x.size(1), dtype=F32, device=x.device)
        arg = torch.empty(K, x.size(1), dtype=I32, device=x.device)
        ops.maxpool_fwd(x, cmptr, cmem, out, arg)
        ctx.n = x.size(0)
        ctx.save_for_backward(arg, cl)
        return out

    @staticmethod
    def backward(ctx, g):
        arg, cl = ctx.saved_tensors
        dx = torch.empty(ctx.n, g.size(1), dtype=F32, device=g.device)
        ops.maxpool_bwd(g.contiguous(), arg, cl, dx)
        return dx, None, None, None, None


class _SegmentMean(torch.autograd.Function):
    """torch_scatter.scatter_mean(x, batch) for sorted ``batch`` given as segment pointers."""

    @staticmethod
    def forward(ctx, x, seg_ptr):
        x = x.contiguous()
        out = torch.empty(seg_ptr.numel() - 1, x.size(1), dtype=F32, device=x.device)
        ops.segment_mean_fwd(x, seg_ptr, out)
        ctx.n = x.size(0)
        ctx.save_for_backward(seg_ptr)
        return out

    @staticmethod
    def backward(ctx, g):
        seg_ptr, = ctx.saved_tensors
        dx = torch.zeros(ctx.n, g.size(1), dtype=F32, device=g.device)
        ops.segment_mean_bwd(g.contiguous(), seg_ptr, dx)
        return dx, None


def aggregate_sum(x, gop):
    return _AggSum.apply(x, gop)


def mean_concat(x, gop, weighted):
    return _MeanConcat.apply(x, gop, weighted)


def linear(X, W, bias=None, Fin=None, Fout=None, groups=1, w_layout=0, relu=False, keep_mask=None, keep_scale=1.0):
    """``w_layout`` 0: W is [groups*Fout, Fin] (nn.Linear.weight); 1: W is [groups*Fin, Fout]."""
    if Fin is None or Fout is None:
        if w_layout == 0:
            Fout, Fin = W.size(0) // groups, W.size(1)
        else:
            Fin, Fout = W.size(0) // groups, W.size(1)
    return _Linear.apply(X, W, bias, Fin, Fout, groups, w_layout, relu, keep_mask, keep_scale)


def cluster_max_pool(x, cmptr, cmem, cl, K):
    return _MaxPool.apply(x, cmptr, cmem, cl, K)


def segment_mean(x, seg_ptr):
    return _SegmentMean.apply(x, seg_ptr)
