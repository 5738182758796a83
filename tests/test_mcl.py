"""Pre-clustering: Markov clustering as ``PreCluster`` runs it (DataSet.py:45-88,
community_pooling.py:142-155).

The shipped fixture stores the 20 cluster vectors the REFERENCE wrote (``clustering/mcl/depth_{0,1}`` of 10
graphs): they pin the oracle restatement of the un-vendored ``markov_clustering`` package (CPU) and the CUDA
kernel ``mcl_graph_kernel`` (GPU), bit for bit.  Random graphs extend the comparison CUDA == oracle."""
import numpy as np
import pytest
import torch

from conftest import FIXTURE


def _fixture_graphs():
    from deeprank_gnn_b200 import hdf5min
    f = hdf5min.File(FIXTURE, 'r')
    out = []
    for mol in f.keys():
        g = f[mol]
        out.append(dict(mol=mol, n=len(g['nodes'][()]), iei=np.array(g['internal_edge_index'][()]).T.copy(),
                        d0=np.array(g['clustering/mcl/depth_0'][()]), d1=np.array(g['clustering/mcl/depth_1'][()])))
    return out


def _pooled_internal_edges(c0, iei):
    """community_pooling of the internal edges (community_pooling.py:208-210): what depth_1 is clustered on."""
    from oracle import pyg_min
    inv, _perm = pyg_min.consecutive_cluster(torch.as_tensor(c0))
    both = torch.from_numpy(np.ascontiguousarray(np.hstack((iei, np.flip(iei, 0))))).long()     # DataSet.py:289-306
    pei, _ = pyg_min.pool_edge(inv, both, torch.ones(both.size(1), 1))
    return pei.numpy(), int(inv.max()) + 1


def test_oracle_mcl_reproduces_the_clusters_stored_by_the_reference():
    from oracle import mcl
    for g in _fixture_graphs():
        c0 = mcl.community_detection_mcl(g['iei'], g['n'])
        assert np.array_equal(c0, g['d0']), g['mol']
        pei, K = _pooled_internal_edges(c0, g['iei'])
        assert K == len(g['d1'])
        assert np.array_equal(mcl.community_detection_mcl(pei, K), g['d1']), g['mol']


def test_dense_iteration_equals_the_sparse_one():
    """The dense float64 form the CUDA kernel implements converges to the same clusters as the sparse package form."""
    from oracle import mcl
    g = _fixture_graphs()[4]
    a = mcl.adjacency(g['iei'], g['n']).toarray()
    m_dense, _it = mcl.run_mcl_dense(a)
    index = np.zeros(g['n'], dtype=np.int64)
    for ic, c in enumerate(mcl.get_clusters(m_dense)):
        index[list(c)] = ic
    assert np.array_equal(index, g['d0'])


@pytest.mark.gpu
def test_cuda_mcl_reproduces_the_clusters_stored_by_the_reference(lib):
    from deeprank_gnn_b200.community_pooling import community_detection, mcl_detection_batch
    gs = _fixture_graphs()
    c0s = mcl_detection_batch([torch.from_numpy(g['iei']) for g in gs], [g['n'] for g in gs])
    for g, c0 in zip(gs, c0s):
        assert np.array_equal(c0.numpy(), g['d0']), g['mol']
    pooled = [_pooled_internal_edges(g['d0'], g['iei']) for g in gs]
    c1s = mcl_detection_batch([torch.from_numpy(p) for p, _K in pooled], [K for _p, K in pooled])
    for g, c1 in zip(gs, c1s):
        assert np.array_equal(c1.numpy(), g['d1']), g['mol']
    # the reference's per-graph entry point, tensors on the device
    one = community_detection(torch.from_numpy(gs[0]['iei']).cuda(), gs[0]['n'], method='mcl')
    assert one.is_cuda and np.array_equal(one.cpu().numpy(), gs[0]['d0'])


@pytest.mark.gpu
def test_cuda_mcl_equals_oracle_on_random_graphs(lib):
    from deeprank_gnn_b200.community_pooling import mcl_detection_batch
    from oracle import mcl
    rng = np.random.default_rng(5)
    eis, ns = [], []
    for n in (2, 7, 33, 64, 150, 257):
        # chain-like graph with random chords, an isolated node, a duplicate edge and a self loop
        src = np.arange(n - 1)
        e = np.stack([src, src + 1])
        k = max(1, n // 3)
        chords = rng.integers(0, max(n - 1, 1), size=(2, k))
        e = np.hstack([e, chords, e[:, :1], np.array([[0], [0]])])
        if n > 10:
            e = e[:, (e[0] != n - 1) & (e[1] != n - 1)]          # the last node has no edge at all
        eis.append(torch.from_numpy(np.ascontiguousarray(e)))
        ns.append(n)
    got = mcl_detection_batch(eis, ns)
    for ei, n, c in zip(eis, ns, got):
        assert np.array_equal(c.numpy(), mcl.community_detection_mcl(ei.numpy(), n)), n


@pytest.mark.gpu
def test_precluster_fills_missing_clusters_on_the_gpu(lib):
    """HDF5DataSet + PreCluster on graphs whose stored clusters are hidden: the GPU kernel recomputes both levels
    and they equal what the reference stored."""
    from deeprank_gnn_b200 import DataSet as ds_mod
    stored = ds_mod.HDF5DataSet(database=FIXTURE, node_feature=['type', 'polarity', 'bsa'], target='irmsd')
    # a clustering method the file does not store: every graph comes without cluster0 / cluster1
    ds = ds_mod.HDF5DataSet(database=FIXTURE, node_feature=['type', 'polarity', 'bsa'], target='irmsd',
                            clustering_method='absent')
    assert getattr(ds.get(0), 'cluster0', None) is None
    ds_mod.PreCluster(ds, 'mcl')
    for i in range(ds.len()):
        a, b = ds.get(i), stored.get(i)
        assert torch.equal(a.cluster0.cpu(), b.cluster0) and torch.equal(a.cluster1.cpu(), b.cluster1), i
