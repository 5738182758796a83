"""Drop-in ``community_pooling`` module: same public names as the reference
(``deeprank_gnn/community_pooling.py``): ``get_preloaded_cluster`` (:25-30),
``community_pooling`` (:161-251), ``community_detection`` (:95-158),
``community_detection_per_batch`` (:33-92).

The runtime half runs on the GPU through the structure pass (one launch pair computes the
cluster offsets, the dense relabelling, the coarsened sorted/unique ``edge_index`` with summed
``edge_attr`` and the pooled ``batch``) plus the cluster max-pool kernel; the detection half
is offline preprocessing and needs the same optional third-party packages as the reference.
"""
import numpy as np
import torch

from . import functional as Fn
from . import ops
from ._lib import DrgnnError
from .data import Batch, Data

I32, I64, F32 = torch.int32, torch.int64, torch.float32


# ------------------------------------------------------------------------------------------
# graph pointers of a batch
# ------------------------------------------------------------------------------------------
def graph_pointers(data):
    """(node_ptr, edge_ptr, B, max_n, max_e) for a ``Batch`` / ``Data`` on the device.  Batches
    collated by ``Batch.from_data_list`` carry them; for hand-made batches they are derived
    from ``batch`` (one host sync)."""
    dev = data.x.device if data.x is not None else data.edge_index.device
    N = data.num_nodes
    E = 0 if data.edge_index is None else data.edge_index.size(1)
    batch = getattr(data, 'batch', None)
    if getattr(data, '_node_ptr', None) is not None and getattr(data, '_edge_ptr', None) is not None:
        return (data._node_ptr.to(dev), data._edge_ptr.to(dev), data._node_ptr.numel() - 1, int(data._max_n),
                int(data._max_e))
    if batch is None:
        node_ptr = torch.tensor([0, N], dtype=I32, device=dev)
        edge_ptr = torch.tensor([0, E], dtype=I32, device=dev)
        return node_ptr, edge_ptr, 1, N, E
    B = int(batch.max()) + 1
    node_ptr = ops.ptr_from_sorted_ids(batch.contiguous(), B)
    egraph = batch[data.edge_index[0]].contiguous() if E else torch.zeros(0, dtype=I64, device=dev)
    edge_ptr = ops.ptr_from_sorted_ids(egraph, B)
    host = torch.stack([node_ptr[1:] - node_ptr[:-1], edge_ptr[1:] - edge_ptr[:-1]]).max(dim=1)[0].tolist()
    return node_ptr, edge_ptr, B, int(host[0]), int(host[1])


def _as_attr(edge_attr):
    if edge_attr is None:
        return None
    edge_attr = edge_attr.to(F32)
    return (edge_attr.unsqueeze(-1) if edge_attr.dim() == 1 else edge_attr).contiguous()


def batch_structure(data, cluster0=None, clusters_are_local=True, use_cluster1=True, edge_index=None,
                    edge_attr='__own__', mirrors=True):
    """Run the structure pass for ``data`` (cached on the object for the default arguments)."""
    default = cluster0 is None and edge_index is None and edge_attr == '__own__' and use_cluster1
    if default and getattr(data, '_structure', None) is not None:
        return data._structure
    node_ptr, edge_ptr, B, max_n, max_e = graph_pointers(data)
    ei = data.edge_index if edge_index is None else edge_index
    if edge_index is not None:        # another edge set over the same nodes (internal edges)
        dev = ei.device
        batch = getattr(data, 'batch', None)
        if batch is None:
            edge_ptr = torch.tensor([0, ei.size(1)], dtype=I32, device=dev)
            max_e = ei.size(1)
        else:
            edge_ptr = ops.ptr_from_sorted_ids(batch[ei[0]].contiguous(), B)
            max_e = int((edge_ptr[1:] - edge_ptr[:-1]).max())
    ea = _as_attr(getattr(data, 'edge_attr', None)) if edge_attr == '__own__' else _as_attr(edge_attr)
    c0 = data.cluster0 if cluster0 is None else cluster0
    if c0 is None:
        raise DrgnnError('no cluster assignment: the batch has no cluster0 (run PreCluster, DataSet.py:45-88)')
    c1 = getattr(data, 'cluster1', None) if use_cluster1 else None
    c1_ptr = None
    if c1 is not None:
        c1_ptr = getattr(data, '_c1_ptr', None)
        if c1_ptr is None:
            # len(cluster1 of graph g) == number of level-0 clusters of g: take the segments from a
            # level-0-only pass
            pre = ops.structure_build(node_ptr, edge_ptr, ei.contiguous(), c0.contiguous(), max_n, max_e,
                                      clusters_are_local=clusters_are_local)
            c1_ptr = pre.kptr0[:B + 1].clone()
        c1_ptr = c1_ptr.to(c0.device)
        c1 = c1.contiguous()
    st = ops.structure_build(node_ptr, edge_ptr, ei.contiguous(), c0.contiguous(), max_n, max_e, c1_ptr=c1_ptr,
                             cluster1=c1, edge_attr=ea, clusters_are_local=clusters_are_local, mirrors=mirrors)
    st.graph_B = B
    if default:
        data._structure = st
    return st


# ------------------------------------------------------------------------------------------
# public API
# ------------------------------------------------------------------------------------------
def get_preloaded_cluster(cluster, batch):
    """Make per-graph cluster ids globally unique, in place (community_pooling.py:25-30):
    ``cluster[batch == g] += sum_{h < g} (max(cluster of h) + 1)``.  One segmented max, one
    scan and one add on the device instead of the reference's Python loop over graphs."""
    if not cluster.is_cuda:
        raise DrgnnError('get_preloaded_cluster runs on CUDA tensors (no CPU fallback)')
    B = int(batch.max()) + 1
    seg = ops.ptr_from_sorted_ids(batch.contiguous(), B)
    return ops.cluster_offset_(cluster, seg)


def community_pooling(cluster, data):
    """Pool a ``Batch`` / ``Data`` over ``cluster`` (community_pooling.py:161-251): dense
    relabelling, cluster max of ``x`` (differentiable), coarsened ``edge_index`` (sorted, unique,
    no self loops) with summed ``edge_attr``, pooled internal edges, cluster mean of ``pos`` /
    ``pos2D`` and pooled ``batch``; ``cluster0`` / ``cluster1`` are carried over."""
    if getattr(data, 'pos', None) is None:
        # the reference raises here too (pos is referenced unconditionally, community_pooling.py:226)
        raise UnboundLocalError("local variable 'pos' referenced before assignment")
    st = batch_structure(data, cluster0=cluster, clusters_are_local=False, use_cluster1=False)
    K0, E1, _ = st.sync_counts()
    x = data.x if data.x.dim() == 2 else data.x.unsqueeze(-1)
    xp = Fn.cluster_max_pool(x.to(F32), st.cmptr0, st.cmem0, st.cl0, K0)
    if data.x.dim() == 1:
        xp = xp.squeeze(-1)
    edge_index = st.edge_index1[:, :E1].clone()
    edge_attr = None if st.ne == 0 else st.edge_attr1[:E1].clone()
    members = Fn.GraphOp(st.cmptr0, st.cmem0, None, None, K0)

    def cluster_mean(t):
        t2 = (t if t.dim() == 2 else t.unsqueeze(-1)).to(F32).contiguous()
        out = torch.empty(K0, t2.size(1), dtype=F32, device=t2.device)
        ops.aggregate(t2, members.rowptr, members.col, out, post_mode=1)
        return out
    pos = cluster_mean(data.pos)
    has_batch = hasattr(data, 'batch')
    if has_batch:
        batch = None if data.batch is None else st.batch1_i64[:K0].clone()
        out = Batch(batch=batch, x=xp, edge_index=edge_index, edge_attr=edge_attr, pos=pos)
        out._num_graphs = st.graph_B
    else:
        out = Data(x=xp, edge_index=edge_index, edge_attr=edge_attr, pos=pos)
        if getattr(data, 'pos2D', None) is not None:
            out.pos2D = cluster_mean(data.pos2D)
    iei = getattr(data, 'internal_edge_index', None)
    if iei is not None:
        sti = batch_structure(data, cluster0=cluster, clusters_are_local=False, use_cluster1=False, edge_index=iei,
                              edge_attr=getattr(data, 'internal_edge_attr', None))
        _, Ei, _ = sti.sync_counts()
        out.internal_edge_index = sti.edge_index1[:, :Ei].clone()
        out.internal_edge_attr = None if sti.ne == 0 else sti.edge_attr1[:Ei].clone()
    if getattr(data, 'cluster0', None) is not None:
        out.cluster0 = data.cluster0
        out.cluster1 = getattr(data, 'cluster1', None)
    return out


# ------------------------------------------------------------------------------------------
# offline community detection (not on the hot path; same optional dependencies as the reference)
# ------------------------------------------------------------------------------------------
def mcl_detection_batch(edge_indices, num_nodes, device=None):
    """Markov clustering of MANY graphs in one launch of the GPU kernel (``drgnn_mcl_cluster``, csrc/mcl.cu: the
    algorithm of ``markov_clustering.run_mcl`` with default parameters + ``get_clusters`` + the labelling of
    community_pooling.py:148-153, float64).  ``edge_indices``: list of ``[2, e_g]`` integer tensors with graph-LOCAL
    node ids (unit weights, undirected), ``num_nodes``: list of node counts.  Returns a list of int64 CPU tensors."""
    from . import ops
    device = torch.device('cuda' if device is None else device)
    node_ptr, edge_ptr, parts = [0], [0], []
    for ei, n in zip(edge_indices, num_nodes):
        ei = torch.as_tensor(ei).long().reshape(2, -1)
        parts.append(ei + node_ptr[-1])
        node_ptr.append(node_ptr[-1] + int(n))
        edge_ptr.append(edge_ptr[-1] + ei.size(1))
    if not parts:
        return []
    ei = torch.cat(parts, dim=1).contiguous().to(device)
    nptr = torch.tensor(node_ptr, dtype=torch.int32, device=device)
    eptr = torch.tensor(edge_ptr, dtype=torch.int32, device=device)
    out = ops.mcl_cluster(ei, nptr, eptr, max(int(n) for n in num_nodes)).cpu()
    return [out[a:b].clone() for a, b in zip(node_ptr[:-1], node_ptr[1:])]


def community_detection(edge_index, num_nodes, edge_attr=None, method='mcl'):
    """Cluster one graph with Markov clustering or Louvain (community_pooling.py:95-158).  ``method='mcl'``
    without edge weights (what ``PreCluster`` asks for, DataSet.py:76,84) runs on the GPU (``mcl_detection_batch``)
    and returns the labels on ``edge_index``'s device; weighted MCL and Louvain keep the reference's optional
    dependencies (networkx plus ``markov_clustering`` / ``python-louvain``)."""
    if method == 'mcl' and edge_attr is None and torch.cuda.is_available():
        dev = edge_index.device if edge_index.is_cuda else None
        return mcl_detection_batch([edge_index], [num_nodes], device=dev)[0].to(edge_index.device)
    import networkx as nx
    g = nx.Graph()
    g.add_nodes_from(range(num_nodes))
    for iedge, (i, j) in enumerate(edge_index.transpose(0, 1).tolist()):
        if edge_attr is None:
            g.add_edge(i, j)
        else:
            g.add_edge(i, j, weight=edge_attr[iedge])
    if method == 'louvain':
        try:
            import community
        except ImportError as e:
            raise ImportError('community detection method "louvain" needs python-louvain') from e
        cluster = community.best_partition(g)
        return torch.tensor([v for _, v in sorted(cluster.items())])
    if method == 'mcl':
        try:
            import markov_clustering as mc
        except ImportError as e:
            raise ImportError('community detection method "mcl" needs markov_clustering; graphs whose clusters '
                              'are already stored in the HDF5 file do not need it') from e
        matrix = nx.to_scipy_sparse_array(g)
        clusters = mc.get_clusters(mc.run_mcl(matrix))
        index = np.zeros(num_nodes).astype('int')
        for ic, c in enumerate(clusters):
            index[list(c)] = ic
        return torch.tensor(index)
    raise ValueError('Clustering method %s not supported' % method)


def community_detection_per_batch(edge_index, batch, num_nodes, edge_attr=None, method='mcl'):
    """Per-graph detection with globally unique ids (community_pooling.py:33-92)."""
    out = torch.zeros(num_nodes, dtype=torch.long)
    offset = 0
    batch_c, ei_c = batch.cpu(), edge_index.cpu()
    for g in range(int(batch_c.max()) + 1):
        nodes = torch.nonzero(batch_c == g).view(-1)
        lo = int(nodes.min())
        mask = (batch_c[ei_c[0]] == g)
        local = community_detection(ei_c[:, mask] - lo, nodes.numel(),
                                    None if edge_attr is None else edge_attr[mask], method)
        out[nodes] = local + offset
        offset += int(local.max()) + 1
    return out.to(edge_index.device)


def community_pooling_host(cluster, data):
    """Host-side (numpy) pooling of ONE graph, used only by ``PreCluster`` to derive the level-1
    clustering input (DataSet.py:82-84): pooled internal edges and node count."""
    uniq, inv = np.unique(np.asarray(cluster), return_inverse=True)
    K = uniq.size

    def pool_edges(ei):
        if ei is None:
            return None
        e = inv[np.asarray(ei)]
        e = e[:, e[0] != e[1]]
        if e.size == 0:
            return torch.zeros(2, 0, dtype=torch.long)
        key = np.unique(e[0].astype(np.int64) * K + e[1])
        return torch.from_numpy(np.stack([key // K, key % K]))
    x = torch.zeros(K, data.x.size(1) if data.x.dim() == 2 else 1)
    out = Data(x=x, edge_index=pool_edges(data.edge_index), pos=None)
    out.internal_edge_index = pool_edges(getattr(data, 'internal_edge_index', None))
    return out
