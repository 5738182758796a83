"""Drop-in ``ginet`` module: ``GINetConvLayer`` and ``GINet`` with the reference's signatures,
attributes and ``state_dict`` names (``deeprank_gnn/ginet.py:22-78, 81-141``), computed by
the sm_100a kernels.

``GINetConvLayer.forward`` is ``z[i] = sum_{e: row[e]=i} alpha_e * fc(x[col[e]])`` where the
reference's alpha is a softmax over a size-1 dimension (ginet.py:62-66), i.e. exactly 1 for
finite inputs: the layer is ``A (X W^T)`` and the parameters of the attention path
(``fc_edge_attr``, ``fc_attention``) receive exactly-zero gradients.  They are kept (same
names, shapes and initialisation) so reference checkpoints load and save unchanged.
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import functional as Fn
from .nets_common import Levels, node_features


def _uniform(size, tensor):
    if tensor is not None:
        bound = 1.0 / (size ** 0.5)
        tensor.data.uniform_(-bound, bound)


class GINetConvLayer(nn.Module):
    def __init__(self, in_channels, out_channels, number_edge_features=1, bias=False):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.fc = nn.Linear(in_channels, out_channels, bias=bias)
        self.fc_edge_attr = nn.Linear(number_edge_features, number_edge_features, bias=bias)
        self.fc_attention = nn.Linear(2 * out_channels + number_edge_features, 1, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):
        size = self.in_channels
        _uniform(size, self.fc.weight)
        _uniform(size, self.fc_attention.weight)
        _uniform(size, self.fc_edge_attr.weight)

    def _dead(self):
        # zero-valued term that gives the dead attention parameters zero (not None) gradients
        return 0.0 * (self.fc_attention.weight.sum() + self.fc_edge_attr.weight.sum())

    def forward(self, x, edge_index, edge_attr=None, graph=None):
        """``graph``: an optional pre-built ``functional.GraphOp`` (network forwards pass the one
        from the structure pass); otherwise CSR/CSC are derived from ``edge_index``."""
        x = x.to(torch.float32)
        if graph is None:
            graph = Fn.GraphOp.from_edge_index(edge_index, x.size(0))
        z = Fn.linear(Fn.aggregate_sum(x, graph), self.fc.weight, None)
        if self.fc.bias is not None:
            # the reference applies fc (with its bias) to x[col] per EDGE before the sum (ginet.py:57,71): a node
            # receives its bias once per incoming edge
            deg = (graph.rowptr[1:] - graph.rowptr[:-1]).to(z.dtype).unsqueeze(1)
            z = z + deg * self.fc.bias
        return z + self._dead()

    def __repr__(self):
        return '{}({}, {})'.format(self.__class__.__name__, self.in_channels, self.out_channels)


class GINet(nn.Module):
    def __init__(self, input_shape, output_shape=1, input_shape_edge=1, hidden=(16, 32)):
        """``hidden`` is an extension (the reference hard-codes 16 / 32, ginet.py:87-93)."""
        super().__init__()
        h1, h2 = hidden
        self.hidden = (h1, h2)
        self.conv1 = GINetConvLayer(input_shape, h1, input_shape_edge)
        self.conv2 = GINetConvLayer(h1, h2, input_shape_edge)
        self.conv1_ext = GINetConvLayer(input_shape, h1, input_shape_edge)
        self.conv2_ext = GINetConvLayer(h1, h2, input_shape_edge)
        self.fc1 = nn.Linear(2 * h2, 4 * h2)
        self.fc2 = nn.Linear(4 * h2, output_shape)
        self.clustering = 'mcl'
        self.dropout = 0.4

    def forward(self, data):
        """Both branches convolve over the same edges and pool with the same clusters
        (ginet.py:101-130), so they share every aggregation / pooling launch: conv1 applies the two
        concatenated weights to one aggregated input, conv2 is a 2-group transform."""
        h1, h2 = self.hidden
        x = node_features(data)
        lv = Levels(data)
        W1 = torch.cat([self.conv1.fc.weight, self.conv1_ext.fc.weight], dim=0)          # [2 h1, F]
        z1 = Fn.linear(Fn.aggregate_sum(x, lv.g0), W1, None, relu=True)                    # [N, 2 h1]
        p1 = lv.pool0(z1)                                                                  # [K0, 2 h1]
        W2 = torch.cat([self.conv2.fc.weight, self.conv2_ext.fc.weight], dim=0)          # [2 h2, h1]
        z2 = Fn.linear(Fn.aggregate_sum(p1, lv.g1), W2, None, Fin=h1, Fout=h2, groups=2, relu=True)   # [K0, 2 h2]
        r = lv.readout(lv.pool1(z2))                                                       # [B, 2 h2] = cat(x, x_ext)
        keep = None
        if self.training and self.dropout > 0:
            keep = torch.empty(r.size(0), self.fc1.out_features, device=r.device).bernoulli_(1.0 - self.dropout)
        h = Fn.linear(r, self.fc1.weight, self.fc1.bias, relu=True, keep_mask=keep,
                      keep_scale=1.0 / (1.0 - self.dropout) if keep is not None else 1.0)
        out = Fn.linear(h, self.fc2.weight, self.fc2.bias)
        dead = self.conv1._dead() + self.conv2._dead() + self.conv1_ext._dead() + self.conv2_ext._dead()
        return out + dead


class GINetInternal(nn.Module):
    """The two-graph GINet of the reference documentation (``docs/tutorial.advanced.rst:126-137``, README "Custom
    GNN"): ``conv1`` / ``conv2`` convolve over the INTERFACE edges (``edge_index``, ``edge_attr``) and
    ``conv1_ext`` / ``conv2_ext`` over the INTERNAL edges (``internal_edge_index``, ``internal_edge_attr``) of the
    same nodes, both pooled with the same clusters; read-outs concatenated -> fc1 -> ReLU -> dropout -> fc2.  (The
    shipped ``ginet.GINet`` feeds ``edge_index`` to both branches, ginet.py:104-126; this is the variant its
    documentation describes.)  Same constructor contract ``(input_shape, output_shape, input_shape_edge)`` and
    ``state_dict`` names as ``GINet``; runs on the generic autograd ops (two structure passes, one per edge set)."""

    def __init__(self, input_shape, output_shape=1, input_shape_edge=1, hidden=(16, 32)):
        super().__init__()
        h1, h2 = hidden
        self.hidden = (h1, h2)
        self.conv1 = GINetConvLayer(input_shape, h1, input_shape_edge)
        self.conv2 = GINetConvLayer(h1, h2, input_shape_edge)
        self.conv1_ext = GINetConvLayer(input_shape, h1, input_shape_edge)
        self.conv2_ext = GINetConvLayer(h1, h2, input_shape_edge)
        self.fc1 = nn.Linear(2 * h2, 4 * h2)
        self.fc2 = nn.Linear(4 * h2, output_shape)
        self.clustering = 'mcl'
        self.dropout = 0.4

    def _branch(self, x, lv, conv1, conv2):
        z1 = Fn.linear(Fn.aggregate_sum(x, lv.g0), conv1.fc.weight, None, relu=True)
        z2 = Fn.linear(Fn.aggregate_sum(lv.pool0(z1), lv.g1), conv2.fc.weight, None, relu=True)
        return lv.readout(lv.pool1(z2))

    def forward(self, data):
        if getattr(data, 'internal_edge_index', None) is None:
            from ._lib import DrgnnError
            raise DrgnnError('GINetInternal needs internal_edge_index (DataSet.py:289-306)')
        x = node_features(data)
        r_int = self._branch(x, Levels(data), self.conv1, self.conv2)
        lv_ext = Levels(data, edge_index=data.internal_edge_index, edge_attr=getattr(data, 'internal_edge_attr', None))
        r_ext = self._branch(x, lv_ext, self.conv1_ext, self.conv2_ext)
        r = torch.cat([r_int, r_ext], dim=1)
        keep = None
        if self.training and self.dropout > 0:
            keep = torch.empty(r.size(0), self.fc1.out_features, device=r.device).bernoulli_(1.0 - self.dropout)
        h = Fn.linear(r, self.fc1.weight, self.fc1.bias, relu=True, keep_mask=keep,
                      keep_scale=1.0 / (1.0 - self.dropout) if keep is not None else 1.0)
        out = Fn.linear(h, self.fc2.weight, self.fc2.bias)
        dead = self.conv1._dead() + self.conv2._dead() + self.conv1_ext._dead() + self.conv2_ext._dead()
        return out + dead
