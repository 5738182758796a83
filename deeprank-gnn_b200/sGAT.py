"""Drop-in ``sGAT`` module: ``sGraphAttentionLayer`` and ``sGAT`` with the reference's signatures
and ``state_dict`` names (``deeprank_gnn/sGAT.py:19-98, 101-138``).

``out_i = (1/max(deg_i,1)) sum_{e: row=i} a_e ([x_i || x_col] W) + b`` is computed as
``[s_i x_i || m_i] W + b`` with ``s_i = sum a_e / d_i`` and ``m_i = sum a_e x_col / d_i``
(one aggregation on N rows, one transform on N rows) instead of a GEMM on E gathered rows.
"""
import torch
import torch.nn as nn
from torch.nn import Parameter

from . import functional as Fn
from ._lib import DrgnnError
from .nets_common import Levels, node_features


def _uniform(size, tensor):
    if tensor is not None:
        bound = 1.0 / (size ** 0.5)
        tensor.data.uniform_(-bound, bound)


class sGraphAttentionLayer(nn.Module):
    def __init__(self, in_channels, out_channels, bias=True, undirected=True):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.undirected = undirected
        self.weight = Parameter(torch.Tensor(2 * in_channels, out_channels))
        if bias:
            self.bias = Parameter(torch.Tensor(out_channels))
        else:
            self.register_parameter('bias', None)
        self.reset_parameters()

    def reset_parameters(self):
        size = 2 * self.in_channels
        _uniform(size, self.weight)
        _uniform(size, self.bias)

    def forward(self, x, edge_index, edge_attr, graph=None, relu=False):
        if not self.undirected:
            raise DrgnnError('undirected=False (second scatter over col, sGAT.py:86-87) is not supported')
        x = x.to(torch.float32)
        if graph is None:
            ea = edge_attr.unsqueeze(-1) if edge_attr.dim() == 1 else edge_attr
            if ea.size(1) != 1:
                raise DrgnnError('sGraphAttentionLayer supports one edge feature')
            graph = Fn.GraphOp.from_edge_index(edge_index, x.size(0), ea[:, 0])
        return Fn.linear(Fn.mean_concat(x, graph, True), self.weight, self.bias, w_layout=1, relu=relu)

    def __repr__(self):
        return '{}({}, {})'.format(self.__class__.__name__, self.in_channels, self.out_channels)


class sGAT(nn.Module):
    def __init__(self, input_shape, output_shape=1, input_shape_edge=None, hidden=(16, 32)):
        super().__init__()
        h1, h2 = hidden
        self.conv1 = sGraphAttentionLayer(input_shape, h1)
        self.conv2 = sGraphAttentionLayer(h1, h2)
        self.fc1 = nn.Linear(h2, 2 * h2)
        self.fc2 = nn.Linear(2 * h2, output_shape)
        self.clustering = 'mcl'

    def forward(self, data):
        x = node_features(data)
        lv = Levels(data)
        if lv.st.ne != 1:
            raise DrgnnError('sGAT needs exactly one edge feature (edge_attr [E,1])')
        z1 = self.conv1(x, None, None, graph=lv.g0, relu=True)
        z2 = self.conv2(lv.pool0(z1), None, None, graph=lv.g1, relu=True)
        r = lv.readout(lv.pool1(z2))
        h = Fn.linear(r, self.fc1.weight, self.fc1.bias, relu=True)
        return Fn.linear(h, self.fc2.weight, self.fc2.bias)
