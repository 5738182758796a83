"""Restatement of the reference model plugins.  TEST INFRASTRUCTURE ONLY.

``GINetConvLayer`` / ``GINet``           <- deeprank_gnn/ginet.py:22-78, 81-141
``sGraphAttentionLayer`` / ``sGAT``      <- deeprank_gnn/sGAT.py:19-98, 101-138
``FoutLayer`` / ``FoutNet``              <- deeprank_gnn/foutnet.py:15-87, 90-125

``LITERAL = True`` keeps the reference's Python loops (``get_preloaded_cluster`` per
graph, ``FoutLayer`` per node).  ``LITERAL = False`` swaps in the loop-free forms
(proven equal in tests/test_oracle.py) so the big configurations finish in seconds.
State-dict names and shapes are the reference's (checked against the shipped
checkpoints in tests/test_checkpoints.py).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn import Parameter

from . import pooling
from .pyg_min import max_pool_x, scatter_mean, scatter_sum, uniform

LITERAL = True


def _offset(cluster, batch):
    if LITERAL:
        return pooling.get_preloaded_cluster(cluster, batch)
    return pooling.get_preloaded_cluster_closed_form(cluster, batch)


# ------------------------------------------------------------------ GINet ---
class GINetConvLayer(nn.Module):
    def __init__(self, in_channels, out_channels, number_edge_features=1, bias=False):
        super().__init__()
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.fc = nn.Linear(in_channels, out_channels, bias=bias)                        # ginet.py:36
        self.fc_edge_attr = nn.Linear(number_edge_features, number_edge_features, bias=bias)
        self.fc_attention = nn.Linear(2 * out_channels + number_edge_features, 1, bias=bias)
        self.reset_parameters()

    def reset_parameters(self):                                                          # ginet.py:43-48
        size = self.in_channels
        uniform(size, self.fc.weight)
        uniform(size, self.fc_attention.weight)
        uniform(size, self.fc_edge_attr.weight)

    def forward(self, x, edge_index, edge_attr):                                         # ginet.py:50-73
        row, col = edge_index
        num_node = len(x)
        edge_attr = edge_attr.unsqueeze(-1) if edge_attr.dim() == 1 else edge_attr
        xcol = self.fc(x[col])
        xrow = self.fc(x[row])
        ed = self.fc_edge_attr(edge_attr)
        alpha = torch.cat([xrow, xcol, ed], dim=1)
        alpha = self.fc_attention(alpha)
        alpha = F.leaky_relu(alpha)
        alpha = F.softmax(alpha, dim=1)          # over a size-1 dim: == 1 for finite inputs
        h = alpha * xcol
        out = torch.zeros(num_node, self.out_channels)
        return scatter_sum(h, row, dim=0, out=out)


class GINet(nn.Module):
    def __init__(self, input_shape, output_shape=1, input_shape_edge=1, hidden=(16, 32)):  # ginet.py:85-97
        # ``hidden`` is NOT in the reference (widths are hard-coded 16/32); it exists so
        # BASELINE cfg4 ("32->64 hidden") can be expressed.  Defaults = reference.
        super().__init__()
        h1, h2 = hidden
        self.conv1 = GINetConvLayer(input_shape, h1, input_shape_edge)
        self.conv2 = GINetConvLayer(h1, h2, input_shape_edge)
        self.conv1_ext = GINetConvLayer(input_shape, h1, input_shape_edge)
        self.conv2_ext = GINetConvLayer(h1, h2, input_shape_edge)
        self.fc1 = nn.Linear(2 * h2, 4 * h2)
        self.fc2 = nn.Linear(4 * h2, output_shape)
        self.clustering = 'mcl'
        self.dropout = 0.4

    def forward(self, data):                                                             # ginet.py:99-141
        act = F.relu
        data_ext = data.clone()

        data.x = act(self.conv1(data.x, data.edge_index, data.edge_attr))
        cluster = _offset(data.cluster0, data.batch)
        data = pooling.community_pooling(cluster, data)
        data.x = act(self.conv2(data.x, data.edge_index, data.edge_attr))
        cluster = _offset(data.cluster1, data.batch)
        x, batch = max_pool_x(cluster, data.x, data.batch)

        data_ext.x = act(self.conv1_ext(data_ext.x, data_ext.edge_index, data_ext.edge_attr))
        cluster = _offset(data_ext.cluster0, data_ext.batch)
        data_ext = pooling.community_pooling(cluster, data_ext)
        data_ext.x = act(self.conv2_ext(data_ext.x, data_ext.edge_index, data_ext.edge_attr))
        cluster = _offset(data_ext.cluster1, data_ext.batch)
        x_ext, batch_ext = max_pool_x(cluster, data_ext.x, data_ext.batch)

        x = scatter_mean(x, batch, dim=0)
        x_ext = scatter_mean(x_ext, batch_ext, dim=0)
        x = torch.cat([x, x_ext], dim=1)
        x = act(self.fc1(x))
        x = F.dropout(x, self.dropout, training=self.training)
        return self.fc2(x)


class GINetInternal(GINet):
    """The two-graph variant of the reference documentation (docs/tutorial.advanced.rst:126-137): identical to
    ``GINet`` except that the ``_ext`` branch convolves over ``internal_edge_index`` / ``internal_edge_attr``."""

    def forward(self, data):
        act = F.relu
        data_ext = data.clone()

        data.x = act(self.conv1(data.x, data.edge_index, data.edge_attr))
        cluster = _offset(data.cluster0, data.batch)
        data = pooling.community_pooling(cluster, data)
        data.x = act(self.conv2(data.x, data.edge_index, data.edge_attr))
        cluster = _offset(data.cluster1, data.batch)
        x, batch = max_pool_x(cluster, data.x, data.batch)

        data_ext.x = act(self.conv1_ext(data_ext.x, data_ext.internal_edge_index, data_ext.internal_edge_attr))
        cluster = _offset(data_ext.cluster0, data_ext.batch)
        data_ext = pooling.community_pooling(cluster, data_ext)
        data_ext.x = act(self.conv2_ext(data_ext.x, data_ext.internal_edge_index, data_ext.internal_edge_attr))
        cluster = _offset(data_ext.cluster1, data_ext.batch)
        x_ext, batch_ext = max_pool_x(cluster, data_ext.x, data_ext.batch)

        x = scatter_mean(x, batch, dim=0)
        x_ext = scatter_mean(x_ext, batch_ext, dim=0)
        x = torch.cat([x, x_ext], dim=1)
        x = act(self.fc1(x))
        x = F.dropout(x, self.dropout, training=self.training)
        return self.fc2(x)


# ------------------------------------------------------------------- sGAT ---
class sGraphAttentionLayer(nn.Module):
    def __init__(self, in_channels, out_channels, bias=True, undirected=True):           # sGAT.py:35-56
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

    def reset_parameters(self):                                                          # sGAT.py:57-60
        size = 2 * self.in_channels
        uniform(size, self.weight)
        uniform(size, self.bias)

    def forward(self, x, edge_index, edge_attr):                                         # sGAT.py:62-93
        row, col = edge_index
        num_node = len(x)
        edge_attr = edge_attr.unsqueeze(-1) if edge_attr.dim() == 1 else edge_attr
        alpha = torch.cat([x[row], x[col]], dim=-1)
        alpha = torch.mm(alpha, self.weight)
        alpha = edge_attr * alpha
        out = torch.zeros(num_node, self.out_channels)
        out = scatter_mean(alpha, row, dim=0, out=out)
        if not self.undirected:
            out = scatter_mean(alpha, col, dim=0, out=out)
        if self.bias is not None:
            out = out + self.bias
        return out


class sGAT(nn.Module):
    def __init__(self, input_shape, output_shape=1, input_shape_edge=None, hidden=(16, 32)):  # sGAT.py:103-112
        super().__init__()
        h1, h2 = hidden
        self.conv1 = sGraphAttentionLayer(input_shape, h1)
        self.conv2 = sGraphAttentionLayer(h1, h2)
        self.fc1 = nn.Linear(h2, 2 * h2)
        self.fc2 = nn.Linear(2 * h2, output_shape)
        self.clustering = 'mcl'

    def forward(self, data):                                                             # sGAT.py:114-138
        act = F.relu
        data.x = act(self.conv1(data.x, data.edge_index, data.edge_attr))
        cluster = _offset(data.cluster0, data.batch)
        data = pooling.community_pooling(cluster, data)
        data.x = act(self.conv2(data.x, data.edge_index, data.edge_attr))
        cluster = _offset(data.cluster1, data.batch)
        x, batch = max_pool_x(cluster, data.x, data.batch)
        x = scatter_mean(x, batch, dim=0)
        x = act(self.fc1(x))
        return self.fc2(x)


class sGAT3(sGAT):
    """The "sGAT 3-layer + community_pooling" throughput variant of BASELINE config 3: the reference
    sGAT (two conv layers) with a third ``sGraphAttentionLayer(h2, h2)`` on the coarsened graph before
    ``max_pool_x``.  Not a reference network - it exists so the fused three-layer path has an oracle."""

    def __init__(self, input_shape, output_shape=1, input_shape_edge=None, hidden=(16, 32)):
        super().__init__(input_shape, output_shape, input_shape_edge, hidden)
        self.conv3 = sGraphAttentionLayer(hidden[1], hidden[1])

    def forward(self, data):
        act = F.relu
        data.x = act(self.conv1(data.x, data.edge_index, data.edge_attr))
        cluster = _offset(data.cluster0, data.batch)
        data = pooling.community_pooling(cluster, data)
        data.x = act(self.conv2(data.x, data.edge_index, data.edge_attr))
        data.x = act(self.conv3(data.x, data.edge_index, data.edge_attr))
        cluster = _offset(data.cluster1, data.batch)
        x, batch = max_pool_x(cluster, data.x, data.batch)
        x = scatter_mean(x, batch, dim=0)
        x = act(self.fc1(x))
        return self.fc2(x)


# ---------------------------------------------------------------- FoutNet ---
class FoutLayer(nn.Module):
    def __init__(self, in_channels, out_channels, bias=True):                            # foutnet.py:29-48
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

    def reset_parameters(self):                                                          # foutnet.py:50-54
        size = self.in_channels
        uniform(size, self.Wc)
        uniform(size, self.Wn)
        uniform(size, self.bias)

    def forward(self, x, edge_index):                                                    # foutnet.py:56-82
        row, col = edge_index
        num_node = len(x)
        alpha = torch.mm(x, self.Wc)
        beta = torch.mm(x, self.Wn)
        if LITERAL:
            gamma = torch.zeros(num_node, self.out_channels)
            for n in range(num_node):                                                    # :71-73
                index = edge_index[:, edge_index[0, :] == n][1, :]
                gamma[n, :] = torch.mean(beta[index, :], dim=0)    # NaN when no neighbour
        else:
            s = torch.zeros(num_node, self.out_channels).index_add_(0, row, beta[col])
            deg = torch.zeros(num_node).index_add_(0, row, torch.ones(row.numel()))
            gamma = s / deg.unsqueeze(-1)                          # 0/0 = NaN, like mean(empty)
        alpha = alpha + gamma
        if self.bias is not None:
            alpha = alpha + self.bias
        return alpha


class FoutNet(nn.Module):
    def __init__(self, input_shape, output_shape=1, input_shape_edge=None, hidden=(16, 32)):  # foutnet.py:92-101
        super().__init__()
        h1, h2 = hidden
        self.conv1 = FoutLayer(input_shape, h1)
        self.conv2 = FoutLayer(h1, h2)
        self.fc1 = nn.Linear(h2, 2 * h2)
        self.fc2 = nn.Linear(2 * h2, output_shape)
        self.clustering = 'mcl'

    def forward(self, data):                                                             # foutnet.py:103-125
        act = F.relu
        data.x = act(self.conv1(data.x, data.edge_index))
        cluster = _offset(data.cluster0, data.batch)
        data = pooling.community_pooling(cluster, data)
        data.x = act(self.conv2(data.x, data.edge_index))
        cluster = _offset(data.cluster1, data.batch)
        x, batch = max_pool_x(cluster, data.x, data.batch)
        x = scatter_mean(x, batch, dim=0)
        x = act(self.fc1(x))
        return self.fc2(x)
