"""Restatement of the third-party primitives on the reference hot path (CPU, torch).

TEST INFRASTRUCTURE ONLY - see ``oracle/__init__.py``.

None of these live under /root/reference; they are the published behaviour of
torch_scatter 2.0.x, torch_sparse 0.6.x and torch_geometric 1.7.x (the versions
the reference CI resolves, ``.github/workflows/build.yml:36-48``).  Call sites in
the reference are cited per function.
"""
import copy
import math

import torch


# --------------------------------------------------------------------------- #
# Data / Batch  (torch_geometric.data.Data / Batch; used at DataSet.py:335,
# community_pooling.py:225,237, NeuralNet.py:153-154 via DataLoader collation)
# --------------------------------------------------------------------------- #
class Data(object):
    """Attribute bag with the handful of PyG ``Data`` behaviours the path uses."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kw):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.y = y
        self.pos = pos
        for k, v in kw.items():
            setattr(self, k, v)

    # PyG exposes only attributes that are not None as ``keys``
    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith('_')]

    def __contains__(self, key):
        return key in self.keys

    def __getitem__(self, key):
        return getattr(self, key)

    @property
    def num_nodes(self):
        return self.x.size(0)

    @property
    def num_features(self):
        return 1 if self.x.dim() == 1 else self.x.size(1)

    def clone(self):
        # PyG Data.clone(): tensors are cloned, everything else deep-copied (ginet.py:101)
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.clone() if torch.is_tensor(v) else copy.deepcopy(v)
        return out

    def to(self, device):
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(device)
        return self


class Batch(Data):
    def __init__(self, batch=None, **kw):
        super().__init__(**kw)
        self.batch = batch

    @property
    def num_graphs(self):
        return int(self.batch.max()) + 1

    @staticmethod
    def from_data_list(data_list):
        """PyG collation: cat along dim 0, except keys containing ``index`` which are
        cat along the last dim and incremented by the cumulative node count; python
        objects are gathered in lists; ``batch`` holds the graph id of each node."""
        keys = data_list[0].keys
        out = Batch()
        for k in keys:
            out.__dict__[k] = []
        out.batch = []
        cum = 0
        for i, d in enumerate(data_list):
            n = d.num_nodes
            for k in keys:
                v = d[k]
                if torch.is_tensor(v) and 'index' in k:
                    v = v + cum
                out.__dict__[k].append(v)
            out.batch.append(torch.full((n,), i, dtype=torch.long))
            cum += n
        for k in keys:
            items = out.__dict__[k]
            if torch.is_tensor(items[0]):
                dim = -1 if 'index' in k else 0
                out.__dict__[k] = torch.cat(items, dim=dim)
        out.batch = torch.cat(out.batch, dim=0)
        return out


# --------------------------------------------------------------------------- #
# torch_scatter
# --------------------------------------------------------------------------- #
def scatter_sum(src, index, dim=0, out=None, dim_size=None):
    """torch_scatter.scatter_sum - ginet.py:71.  CPU: sequential in edge order."""
    assert dim == 0
    if out is None:
        if dim_size is None:
            dim_size = int(index.max()) + 1 if index.numel() > 0 else 0
        out = torch.zeros((dim_size,) + tuple(src.shape[1:]), dtype=src.dtype)
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    return out.scatter_add_(0, idx, src)


def scatter_mean(src, index, dim=0, out=None, dim_size=None):
    """torch_scatter.scatter_mean - sGAT.py:81, ginet.py:133-134, foutnet.py:120,
    community_pooling.py:214.  sum / clamp(count, min=1); with ``out=`` given the sum
    is accumulated into ``out`` and the whole of ``out`` is then divided."""
    out = scatter_sum(src, index, dim, out, dim_size)
    dim_size = out.size(0)
    ones = torch.ones(index.size(0), dtype=src.dtype)
    count = torch.zeros(dim_size, dtype=src.dtype).scatter_add_(0, index, ones)
    count = count.clamp_(min=1).view(-1, *([1] * (out.dim() - 1)))
    # torch_scatter does out.div_(count) in place for floating tensors; a functional
    # divide keeps autograd happy and is numerically identical
    return out / count


class _ScatterMax(torch.autograd.Function):
    """torch_scatter.scatter_max - community_pooling.py:201 and PyG max_pool_x.

    Published CPU algorithm (torch_scatter 2.0.x ``scatter_cpu`` + ``Reducer<MAX>``): ``out`` is
    filled with ``numeric_limits::lowest()``, ``arg`` with ``src.size(0)``; entries are visited
    in index order and replace the running value only on a strict ``>`` (so the FIRST
    occurrence of the maximum wins and a NaN never wins); afterwards entries still equal to
    ``lowest()`` (empty segments) are set to 0.  Backward: the gradient is routed only to
    ``arg`` (no tie splitting)."""

    @staticmethod
    def forward(ctx, src, index, dim_size):
        n, c = src.shape
        lowest = torch.finfo(src.dtype).min
        idx = index.view(-1, 1).expand(n, c)
        cand_val = torch.where(torch.isnan(src), torch.full_like(src, lowest), src)
        out = torch.full((dim_size, c), lowest, dtype=src.dtype)
        out = out.scatter_reduce(0, idx, cand_val, reduce='amax', include_self=True)
        is_max = (cand_val == out[index]) & (cand_val > lowest)
        pos = torch.arange(n).view(-1, 1).expand(n, c)
        cand = torch.where(is_max, pos, torch.full_like(pos, n))
        arg = torch.full((dim_size, c), n, dtype=torch.long).scatter_reduce(
            0, idx, cand, reduce='amin', include_self=True)
        out = torch.where(out == lowest, torch.zeros_like(out), out)
        ctx.save_for_backward(arg)
        ctx.n = n
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, g, _garg):
        arg, = ctx.saved_tensors
        n = ctx.n
        gs = torch.zeros((n + 1, g.size(1)), dtype=g.dtype)
        gs.scatter_(0, arg, g)       # each (segment, channel) has at most one winner row
        return gs[:n], None, None


def scatter_max(src, index, dim=0, dim_size=None):
    assert dim == 0
    if dim_size is None:
        dim_size = int(index.max()) + 1
    squeeze = src.dim() == 1
    if squeeze:
        src = src.unsqueeze(-1)
    out, arg = _ScatterMax.apply(src, index, dim_size)
    if squeeze:
        out, arg = out.squeeze(-1), arg.squeeze(-1)
    return out, arg


# --------------------------------------------------------------------------- #
# torch_geometric pooling helpers
# --------------------------------------------------------------------------- #
def consecutive_cluster(src):
    """torch_geometric.nn.pool.consecutive.consecutive_cluster - community_pooling.py:197."""
    unique, inv = torch.unique(src, sorted=True, return_inverse=True)
    perm = torch.arange(inv.size(0), dtype=inv.dtype)
    perm = inv.new_empty(unique.size(0)).scatter_(0, inv, perm)
    return inv, perm


def remove_self_loops(edge_index, edge_attr=None):
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    return edge_index, (None if edge_attr is None else edge_attr[mask])


def coalesce(index, value, m, n):
    """torch_sparse.coalesce (op='add'): sort by row*n+col, unique keys, sum values of
    duplicates.  Duplicates are summed in ascending original edge order (stable sort)."""
    row, col = index
    key = row * n + col
    perm = torch.argsort(key, stable=True)
    key = key[perm]
    uniq, inv = torch.unique_consecutive(key, return_inverse=True)
    out_index = torch.stack([torch.div(uniq, n, rounding_mode='floor'), uniq % n], dim=0)
    if value is not None:
        v = value[perm]
        out_val = torch.zeros((uniq.numel(),) + tuple(v.shape[1:]), dtype=v.dtype)
        out_val.index_add_(0, inv, v)
        value = out_val
    return out_index, value


def pool_edge(cluster, edge_index, edge_attr=None):
    """torch_geometric.nn.pool.pool.pool_edge - community_pooling.py:204-205, 209-210."""
    num_nodes = cluster.size(0)
    edge_index = cluster[edge_index.view(-1)].view(2, -1)
    edge_index, edge_attr = remove_self_loops(edge_index, edge_attr)
    if edge_index.numel() > 0:
        edge_index, edge_attr = coalesce(edge_index, edge_attr, num_nodes, num_nodes)
    return edge_index, edge_attr


def pool_batch(perm, batch):
    return batch[perm]


def max_pool_x(cluster, x, batch):
    """torch_geometric.nn.max_pool_x - ginet.py:114,129; sGAT.py:130; foutnet.py:117."""
    cluster, perm = consecutive_cluster(cluster)
    x, _ = scatter_max(x, cluster, dim=0)
    batch = pool_batch(perm, batch)
    return x, batch


def uniform(size, tensor):
    """torch_geometric.nn.inits.uniform - ginet.py:46-48, sGAT.py:59-60, foutnet.py:52-54."""
    if tensor is not None:
        bound = 1.0 / math.sqrt(size)
        tensor.data.uniform_(-bound, bound)
