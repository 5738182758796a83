"""Restatement of ``deeprank_gnn/community_pooling.py`` (runtime half).  TEST INFRASTRUCTURE ONLY.

``get_preloaded_cluster``  <- community_pooling.py:25-30
``community_pooling``      <- community_pooling.py:161-251
"""
import torch

from .pyg_min import (Batch, Data, consecutive_cluster, pool_batch, pool_edge,
                      scatter_max, scatter_mean)


def get_preloaded_cluster(cluster, batch):
    """Literal loop of community_pooling.py:25-30 (in place, cumulative)."""
    nbatch = int(torch.max(batch)) + 1
    for ib in range(1, nbatch):
        cluster[batch == ib] += torch.max(cluster[batch == ib - 1]) + 1
    return cluster


def get_preloaded_cluster_closed_form(cluster, batch):
    """Loop-free equivalent: c[i] = local[i] + sum_{g < batch[i]} (max_local_g + 1).
    Proven equal to the literal loop in tests/test_oracle.py; used where the literal
    O(B*N) loop would take seconds (cfg4 / cfg5 CPU baselines)."""
    nb = int(batch.max()) + 1
    mx = torch.full((nb,), torch.iinfo(torch.int64).min, dtype=torch.int64)
    mx = mx.scatter_reduce(0, batch, cluster, reduce='amax', include_self=True)
    off = torch.cumsum(mx + 1, 0) - (mx + 1)
    cluster += off[batch]
    return cluster


def community_pooling(cluster, data, offset_fn=None):
    """community_pooling.py:161-251."""
    has_internal_edges = hasattr(data, 'internal_edge_index') and data.internal_edge_index is not None
    has_pos2D = hasattr(data, 'pos2D') and data.pos2D is not None
    has_pos = hasattr(data, 'pos') and data.pos is not None
    has_cluster = hasattr(data, 'cluster0') and data.cluster0 is not None

    cluster, perm = consecutive_cluster(cluster)                      # :197

    x, _ = scatter_max(data.x, cluster, dim=0)                        # :201

    edge_index, edge_attr = pool_edge(cluster, data.edge_index, data.edge_attr)   # :204-205

    if has_internal_edges:                                            # :208-210
        internal_edge_index, internal_edge_attr = pool_edge(
            cluster, data.internal_edge_index, data.internal_edge_attr)

    if has_pos:                                                       # :213-216
        pos = scatter_mean(data.pos, cluster, dim=0)
    else:
        # the reference raises UnboundLocalError at :226 when pos is missing
        raise UnboundLocalError("local variable 'pos' referenced before assignment")
    if has_pos2D:
        pos2D = scatter_mean(data.pos2D, cluster, dim=0)

    if has_cluster:
        c0, c1 = data.cluster0, data.cluster1

    if hasattr(data, 'batch'):                                        # :222-234
        batch = None if data.batch is None else pool_batch(perm, data.batch)
        out = Batch(batch=batch, x=x, edge_index=edge_index, edge_attr=edge_attr, pos=pos)
    else:                                                             # :236-249
        out = Data(x=x, edge_index=edge_index, edge_attr=edge_attr, pos=pos)
        if has_pos2D:
            out.pos2D = pos2D
    if has_internal_edges:
        out.internal_edge_index = internal_edge_index
        out.internal_edge_attr = internal_edge_attr
    if has_cluster:
        out.cluster0 = c0
        out.cluster1 = c1
    return out
