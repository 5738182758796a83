"""Shared helpers of the parity tests: product records -> oracle records, oracle structure."""
import copy

import torch

import oracle
from oracle import pyg_min, pooling


def to_oracle_data(d):
    """Product ``Data`` -> oracle ``Data`` (deep copies of the tensors)."""
    kw = {}
    for k in d.keys:
        v = d[k]
        kw[k] = v.clone() if torch.is_tensor(v) else copy.deepcopy(v)
    return pyg_min.Data(**kw)


def to_oracle_batch(graphs):
    return pyg_min.Batch.from_data_list([to_oracle_data(g) for g in graphs])


def oracle_structure(graphs):
    """What the reference computes, integer side, for one mini-batch: offset + relabelled
    level-0 clusters, pooled edges / attrs / batch, offset + relabelled level-1 clusters."""
    b = to_oracle_batch(graphs)
    c0 = pooling.get_preloaded_cluster(b.cluster0.clone(), b.batch)
    inv0, perm0 = pyg_min.consecutive_cluster(c0)
    ei1, ea1 = pyg_min.pool_edge(inv0, b.edge_index, b.edge_attr)
    batch1 = pyg_min.pool_batch(perm0, b.batch)
    c1 = pooling.get_preloaded_cluster(b.cluster1.clone(), batch1)
    inv1, perm1 = pyg_min.consecutive_cluster(c1)
    batch2 = pyg_min.pool_batch(perm1, batch1)
    return dict(batch=b, cl0=inv0, edge_index1=ei1, edge_attr1=ea1, batch1=batch1, cl1=inv1, batch2=batch2,
                K0=int(inv0.max()) + 1, E1=ei1.size(1), K1=int(inv1.max()) + 1)
