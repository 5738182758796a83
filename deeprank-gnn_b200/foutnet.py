"""Drop-in ``foutnet`` module: ``FoutLayer`` and ``FoutNet`` with the reference's signatures and
``state_dict`` names (``deeprank_gnn/foutnet.py:15-87, 90-125``).

``out_i = x_i Wc + mean_{e: row=i}(x_col Wn) + b`` is ``[x_i || mean_i(x_col)] [Wc; Wn] + b``:
one CSR aggregation replaces the reference's Python loop over every node (foutnet.py:71-73).
A node without neighbour yields a NaN row exactly like ``torch.mean`` of an empty selection.
"""
import torch
import torch.nn as nn
from torch.nn import Parameter

from . import functional as Fn
from .nets_common import Levels, node_features


def _uniform(size, tensor):
    if tensor is not None:
        bound = 1.0 / (size ** 0.5)
        tensor.data.uniform_(-bound, bound)


class FoutLayer(nn.Module):
    def __init__(self, in_channels, out_channels, bias=True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.Wc = Parameter(torch.Tensor(in_channels, out_channels))
        self.Wn = Parameter(torch.Tensor(in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        size = self.in_channels
        _uniform(size, self.Wc)
        _uniform(size, self.Wn)
        _uniform(size, self.bias)

    def forward(self, x, edge_index, graph=None, relu=False):
        x = x.to(torch.float32)
        if graph is None:
            graph = Fn.GraphOp.from_edge_index(edge_index, x.size(0))
        W = torch.cat([self.Wc, self.Wn], dim=0)                       # [2 Fin, Fout]
        return Fn.linear(Fn.mean_concat(x, graph, False), W, self.bias, w_layout=1, relu=relu)

    def __repr__(self):
        return '{}({}, {})'.format(self.__class__.__name__, self.in_channels, self.out_channels)


class FoutNet(nn.Module):
    def __init__(self, input_shape, output_shape=1, input_shape_edge=None, hidden=(16, 32)):
        super().__init__()
        h1, h2 = hidden
        self.conv1 = FoutLayer(input_shape, h1)
        self.conv2 = FoutLayer(h1, h2)
        self.fc1 = nn.Linear(h2, 2 * h2)
        self.fc2 = nn.Linear(2 * h2, output_shape)
        self.clustering = 'mcl'

    def forward(self, data):
        x = node_features(data)
        lv = Levels(data)
        z1 = self.conv1(x, None, graph=lv.g0, relu=True)
        z2 = self.conv2(lv.pool0(z1), None, graph=lv.g1, relu=True)
        r = lv.readout(lv.pool1(z2))
        h = Fn.linear(r, self.fc1.weight, self.fc1.bias, relu=True)
        return Fn.linear(h, self.fc2.weight, self.fc2.bias)
