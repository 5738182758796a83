"""Shared forward skeleton of the three reference networks (conv -> cluster offset ->
community_pooling -> conv -> cluster offset -> max_pool_x -> scatter_mean -> fc1 -> fc2;
ginet.py:99-141, sGAT.py:114-138, foutnet.py:103-125) on top of the structure pass."""
import torch

from . import functional as Fn
from .community_pooling import batch_structure


class Levels(object):
    """The two graphs a forward pass convolves over and the two pooling maps, as views of one
    ``Structure`` (ONE host sync for the live sizes K0, E1, K1)."""

    def __init__(self, data, edge_index=None, edge_attr='__own__'):
        """``edge_index`` / ``edge_attr``: another edge set over the same nodes and clusters (the internal edges
        of the two-graph GINet of docs/tutorial.advanced.rst:126-137)."""
        st = batch_structure(data, mirrors=False if getattr(data, '_no_mirrors', False) else True,
                             edge_index=edge_index, edge_attr=edge_attr)
        self.st = st
        K0, E1, K1 = st.sync_counts()
        self.K0, self.E1, self.K1 = K0, E1, K1
        N, B = st.N, st.graph_B
        w = st.ne > 0
        self.g0 = Fn.GraphOp(st.rowptr0[:N + 1], st.col0, st.cscptr0[:N + 1], st.cscrow0, N,
                             st.w0csr if w else None, st.w0csc if w else None)
        self.g1 = Fn.GraphOp(st.rowptr1[:K0 + 1], st.col1, st.cscptr1[:K0 + 1], st.cscrow1, K0,
                             st.edge_attr1[:, 0].contiguous() if w else None, st.w1csc if w else None)
        self.B = B

    def pool0(self, x):
        st = self.st
        return Fn.cluster_max_pool(x, st.cmptr0, st.cmem0, st.cl0, self.K0)

    def pool1(self, x):
        st = self.st
        return Fn.cluster_max_pool(x, st.cmptr1, st.cmem1, st.cl1, self.K1)

    def readout(self, x):
        return Fn.segment_mean(x, self.st.kptr1[:self.B + 1])


def node_features(data):
    x = data.x
    if not x.is_cuda:
        from ._lib import DrgnnError
        raise DrgnnError('the networks run on CUDA only: move the batch with data.to("cuda") (no CPU fallback)')
    x = x if x.dim() == 2 else x.unsqueeze(-1)
    return x.to(torch.float32).contiguous()
