"""GPU parity of every C-ABI kernel against the CPU oracle / plain torch fp32 (``-m gpu``).

Integer outputs (relabelled clusters, pooled edge_index, pooled batch vectors, CSR forms)
must be bit-exact; floating outputs within 1e-4 (BASELINE north_star), usually far tighter.
"""
import numpy as np
import pytest
import torch

from helpers import oracle_structure
from oracle import pyg_min

pytestmark = pytest.mark.gpu

TOL = 1e-4


def _dev():
    return torch.device('cuda:0')


def _graph_sets():
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.DataSet import HDF5DataSet
    from conftest import FIXTURE
    ds = HDF5DataSet(database=FIXTURE, node_feature=['type', 'polarity', 'bsa'], target='irmsd')
    fixture = [ds.get(i) for i in range(ds.len())]
    sets = {
        'fixture10': fixture,
        'fixture1': fixture[:1],
        'cfg2x8': synthetic.make_graphs('cfg2', count=8, seed=3),
        'mixed': synthetic.make_graphs(dict(nodes=(50, 1000), edges_per_node=8, feat=8), count=6, seed=11),
        'cfg4x3': synthetic.make_graphs('cfg4', count=3, seed=5),
    }
    return sets


def _build(graphs, idx_dtype=torch.int64, mirrors=True, with_attr=True):
    from deeprank_gnn_b200 import ops
    from deeprank_gnn_b200.data import Batch
    dev = _dev()
    b = Batch.from_data_list(graphs)
    ei = b.edge_index.to(idx_dtype).to(dev)
    c0 = b.cluster0.to(idx_dtype).to(dev)
    c1 = b.cluster1.to(idx_dtype).to(dev)
    ea = b.edge_attr.to(dev) if with_attr else None
    st = ops.structure_build(b._node_ptr.to(dev), b._edge_ptr.to(dev), ei, c0, b._max_n, b._max_e,
                             c1_ptr=b._c1_ptr.to(dev), cluster1=c1, edge_attr=ea, mirrors=mirrors)
    return b, st


@pytest.mark.parametrize('name', ['fixture10', 'fixture1', 'cfg2x8', 'mixed', 'cfg4x3'])
@pytest.mark.parametrize('idx', [torch.int64, torch.int32])
def test_structure_matches_reference_pooling(lib, name, idx):
    graphs = _graph_sets()[name]
    ref = oracle_structure(graphs)
    b, st = _build(graphs, idx)
    K0, E1, K1 = st.sync_counts()
    assert (K0, E1, K1) == (ref['K0'], ref['E1'], ref['K1'])
    assert torch.equal(st.cl0_i64.cpu(), ref['cl0'])
    assert torch.equal(st.cl0.cpu().long(), ref['cl0'])
    assert torch.equal(st.batch1_i64[:K0].cpu(), ref['batch1'])
    assert torch.equal(st.edge_index1[:, :E1].cpu(), ref['edge_index1'])
    assert torch.allclose(st.edge_attr1[:E1].cpu(), ref['edge_attr1'], rtol=1e-6, atol=1e-6)
    assert torch.equal(st.cl1.cpu().long(), ref['cl1'])
    assert torch.equal(st.batch2_i64[:K1].cpu(), ref['batch2'])
    # CSR of the level-0 graph: rows sorted by destination, ascending original edge id inside a row
    row, col = b.edge_index
    order = torch.from_numpy(np.argsort(row.numpy(), kind='stable'))
    assert torch.equal(st.eid0.cpu().long(), order)
    assert torch.equal(st.col0.cpu().long(), col[order])
    deg = torch.bincount(row, minlength=b.x.size(0))
    assert torch.equal(st.rowptr0.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), deg.cumsum(0)]))
    # CSC
    order_c = torch.from_numpy(np.argsort(col.numpy(), kind='stable'))
    assert torch.equal(st.csceid0.cpu().long(), order_c)
    assert torch.equal(st.cscrow0.cpu().long(), row[order_c])
    assert torch.allclose(st.w0csr.cpu(), b.edge_attr[order, 0])
    assert torch.allclose(st.w0csc.cpu(), b.edge_attr[order_c, 0])
    # pooled CSR / CSC and weights
    r1, c1 = ref['edge_index1']
    deg1 = torch.bincount(r1, minlength=K0)
    assert torch.equal(st.rowptr1[:K0 + 1].cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long), deg1.cumsum(0)]))
    assert torch.equal(st.col1[:E1].cpu().long(), c1)
    oc1 = torch.from_numpy(np.argsort(c1.numpy(), kind='stable'))
    assert torch.equal(st.csceid1[:E1].cpu().long(), oc1)
    assert torch.equal(st.cscrow1[:E1].cpu().long(), r1[oc1])
    assert torch.allclose(st.w1csc[:E1].cpu(), ref['edge_attr1'][oc1, 0], rtol=1e-6, atol=1e-6)
    # members-of-cluster CSR: ascending node id inside a cluster
    cm = st.cmptr0[:K0 + 1].cpu().long()
    mem = st.cmem0.cpu().long()
    srt = torch.from_numpy(np.argsort(ref['cl0'].numpy(), kind='stable'))
    assert torch.equal(mem, srt)
    assert torch.equal(cm, torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(ref['cl0'], minlength=K0).cumsum(0)]))
    srt1 = torch.from_numpy(np.argsort(ref['cl1'].numpy(), kind='stable'))
    assert torch.equal(st.cmem1[:K0].cpu().long(), srt1)
    # per-graph pointers
    assert torch.equal(st.kptr0.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long),
                                                         torch.bincount(ref['batch1'], minlength=st.B).cumsum(0)]))
    assert torch.equal(st.kptr1.cpu().long(), torch.cat([torch.zeros(1, dtype=torch.long),
                                                         torch.bincount(ref['batch2'], minlength=st.B).cumsum(0)]))


def _blob_views(blob, g, n0, e0, n, m):
    """The arrays of graph g inside a structure blob buffer (host tensor), cut to their valid lengths."""
    off = 48 * g + 12 * n0 + 4 * e0
    h = blob[off:off + 32].tolist()
    K, E1, K1 = h[2], h[3], h[4]
    o = off + 32
    out = {'header': h[:6]}
    for name, cap, valid in (('rp0', n + 1, n + 1), ('col0', m, m), ('rp1', n + 1, K + 1), ('col1', m, E1),
                             ('cmp0', n + 1, K + 1), ('cmem0', n, n), ('cl0', n, n), ('cmp1', n + 1, K1 + 1),
                             ('cmem1', n, K), ('cl1', n, K), ('cscp1', n + 1, K + 1), ('cscr1', m, E1)):
        out[name] = blob[o:o + valid]
        o += cap
    return out


@pytest.mark.parametrize('name', ['fixture10', 'fixture1', 'cfg2x8', 'small_mixed'])
@pytest.mark.parametrize('idx', [torch.int64, torch.int32])
def test_blob_structure_pass_equals_full_pass(lib, name, idx):
    """The one-launch bitmap kernel (``drgnn_structure_blob``) against the full structure pass: the
    per-graph blobs are identical word for word, and (graph-local + offsets) equal to the global
    arrays that are themselves checked against the reference pooling above.  Bit-exact."""
    from deeprank_gnn_b200 import ops, synthetic
    from deeprank_gnn_b200.data import Batch
    sets = _graph_sets()
    sets['small_mixed'] = synthetic.make_graphs(dict(nodes=(5, 260), edges_per_node=6, feat=8), count=9, seed=17)
    graphs = sets[name]
    dev = _dev()
    b, st = _build(graphs, idx, mirrors=False, with_attr=False)
    K0, E1tot, K1tot = st.sync_counts()
    assert ops.structure_blob_fits(b._max_n, b._max_e)
    sb = ops.structure_blob(b._node_ptr.to(dev), b._edge_ptr.to(dev), b.edge_index.to(idx).to(dev),
                            b.cluster0.to(idx).to(dev), b._max_n, b._max_e, b._c1_ptr.to(dev),
                            b.cluster1.to(idx).to(dev))
    sb.sync_counts()
    if idx == torch.int32:
        # compact feeder form: uint16 graph-LOCAL edge ids (what PackedBatch(idx16=True) ships), int32 cluster ids
        cnt = (b._edge_ptr[1:] - b._edge_ptr[:-1]).long()
        first = torch.repeat_interleave(b._node_ptr[:-1].long(), cnt)
        ei16 = (b.edge_index - first.unsqueeze(0)).to(torch.int16).to(dev)
        s16 = ops.structure_blob(b._node_ptr.to(dev), b._edge_ptr.to(dev), ei16, b.cluster0.to(idx).to(dev),
                                 b._max_n, b._max_e, b._c1_ptr.to(dev), b.cluster1.to(idx).to(dev))
        s16.sync_counts()
        nw = 48 * len(graphs) + 12 * b.x.size(0) + 4 * b.edge_index.size(1)
        assert torch.equal(s16.blob[:nw].cpu(), sb.blob[:nw].cpu())
        f16 = ops.structure_build(b._node_ptr.to(dev), b._edge_ptr.to(dev), ei16, b.cluster0.to(idx).to(dev), b._max_n,
                                  b._max_e, c1_ptr=b._c1_ptr.to(dev), cluster1=b.cluster1.to(idx).to(dev))
        assert f16.sync_counts() == (K0, E1tot, K1tot)
        assert torch.equal(f16.col0.cpu(), st.col0.cpu()) and torch.equal(f16.col1[:E1tot].cpu(), st.col1[:E1tot].cpu())
    full, lean = st.blob.cpu(), sb.blob.cpu()
    nptr, eptr = b._node_ptr.tolist(), b._edge_ptr.tolist()
    kptr0, kptr1 = st.kptr0.cpu().tolist(), st.kptr1.cpu().tolist()
    rowptr1, cscptr1 = st.rowptr1.cpu(), st.cscptr1.cpu()
    k_seen = e1_seen = 0
    for g in range(len(graphs)):
        n0, e0 = nptr[g], eptr[g]
        n, m = nptr[g + 1] - n0, eptr[g + 1] - e0
        vf, vl = _blob_views(full, g, n0, e0, n, m), _blob_views(lean, g, n0, e0, n, m)
        assert vl['header'] == vf['header'] and vl['header'][5] == 1 and vl['header'][:2] == [n, m]
        for key in vf:
            if key != 'header':
                assert torch.equal(vl[key], vf[key]), (g, key)
        K, E1, K1 = vl['header'][2:5]
        k0, q0 = kptr0[g], kptr1[g]
        assert (K, K1) == (kptr0[g + 1] - k0, kptr1[g + 1] - q0)
        e10 = int(rowptr1[k0])
        assert E1 == int(rowptr1[k0 + K]) - e10
        # graph-local blob == global arrays of the full pass
        assert torch.equal(vl['rp0'] + e0, st.rowptr0[n0:n0 + n + 1].cpu())
        assert torch.equal(vl['col0'] + n0, st.col0[e0:e0 + m].cpu())
        assert torch.equal(vl['cl0'] + k0, st.cl0[n0:n0 + n].cpu())
        assert torch.equal(vl['cmem0'] + n0, st.cmem0[n0:n0 + n].cpu())
        assert torch.equal(vl['cmp0'] + n0, st.cmptr0[k0:k0 + K + 1].cpu())
        assert torch.equal(vl['rp1'] + e10, rowptr1[k0:k0 + K + 1])
        assert torch.equal(vl['col1'] + k0, st.col1[e10:e10 + E1].cpu())
        assert torch.equal(vl['cscp1'] + e10, cscptr1[k0:k0 + K + 1])
        assert torch.equal(vl['cscr1'] + k0, st.cscrow1[e10:e10 + E1].cpu())
        assert torch.equal(vl['cl1'] + q0, st.cl1[k0:k0 + K].cpu())
        assert torch.equal(vl['cmem1'] + k0, st.cmem1[k0:k0 + K].cpu())
        assert torch.equal(vl['cmp1'] + k0, st.cmptr1[q0:q0 + K1 + 1].cpu())
        k_seen += K
        e1_seen += E1
    assert (k_seen, e1_seen) == (K0, E1tot)


def _numpy_blob(n, row, col, c0, c1):
    """The blob arrays of ONE graph from first principles (graph-local ids): consecutive_cluster = rank among
    the sorted unique ids; CSR = stable sort by destination; pool_edge + coalesce = sorted unique (row, col)
    pairs of the relabelled edges without self loops; member lists = stable sort by cluster."""
    def rank(ids):
        u = np.unique(ids)
        return np.searchsorted(u, ids), len(u)

    def ptr(keys, k):
        return np.concatenate([[0], np.cumsum(np.bincount(keys, minlength=k))])
    d0, K = rank(c0)
    d1, K1 = rank(c1)
    order = np.argsort(row, kind='stable')
    out = {'rp0': ptr(row, n), 'col0': col[order], 'cl0': d0, 'cmp0': ptr(d0, K), 'cmem0': np.argsort(d0, kind='stable'),
           'cl1': d1, 'cmp1': ptr(d1, K1), 'cmem1': np.argsort(d1, kind='stable')}
    pr, pc = d0[row], d0[col]
    keep = pr != pc
    pairs = np.unique(np.stack([pr[keep], pc[keep]], axis=1), axis=0) if keep.any() else np.zeros((0, 2), dtype=np.int64)
    out['rp1'], out['col1'] = ptr(pairs[:, 0], K), pairs[:, 1]
    byc = pairs[np.lexsort((pairs[:, 0], pairs[:, 1]))] if len(pairs) else pairs
    out['cscp1'], out['cscr1'] = ptr(byc[:, 1], K), byc[:, 0]
    return out, K, len(pairs), K1


@pytest.mark.parametrize('name', ['fixture10', 'cfg2x8', 'small_mixed'])
def test_blob_structure_pass_matches_first_principles(lib, name):
    """``drgnn_structure_blob`` against an independent numpy construction of every list (not against another
    kernel): bit-exact, on the shipped fixture (real MCL clusters with id gaps) and on synthetic graphs."""
    from deeprank_gnn_b200 import ops, synthetic
    from deeprank_gnn_b200.data import Batch
    sets = _graph_sets()
    sets['small_mixed'] = synthetic.make_graphs(dict(nodes=(5, 260), edges_per_node=6, feat=8), count=9, seed=17)
    graphs = sets[name]
    dev = _dev()
    b = Batch.from_data_list(graphs)
    sb = ops.structure_blob(b._node_ptr.to(dev), b._edge_ptr.to(dev), b.edge_index.to(dev), b.cluster0.to(dev),
                            b._max_n, b._max_e, b._c1_ptr.to(dev), b.cluster1.to(dev))
    sb.sync_counts()
    blob = sb.blob.cpu()
    nptr, eptr, cptr = b._node_ptr.tolist(), b._edge_ptr.tolist(), b._c1_ptr.tolist()
    for g in range(len(graphs)):
        n0, e0 = nptr[g], eptr[g]
        n, m = nptr[g + 1] - n0, eptr[g + 1] - e0
        ei = b.edge_index[:, e0:e0 + m].numpy() - n0
        ref, K, E1, K1 = _numpy_blob(n, ei[0], ei[1], b.cluster0[n0:n0 + n].numpy(), b.cluster1[cptr[g]:cptr[g + 1]].numpy())
        v = _blob_views(blob, g, n0, e0, n, m)
        assert v['header'] == [n, m, K, E1, K1, 1], (g, v['header'])
        for key, arr in ref.items():
            assert np.array_equal(v[key].numpy(), arr), (g, key)


def test_blob_structure_pass_flags_invalid_input(lib):
    from deeprank_gnn_b200 import ops
    from deeprank_gnn_b200._lib import DrgnnError
    dev = _dev()
    node_ptr = torch.tensor([0, 3, 6], dtype=torch.int32, device=dev)
    edge_ptr = torch.tensor([0, 2, 4], dtype=torch.int32, device=dev)
    c1_ptr = torch.tensor([0, 2, 4], dtype=torch.int32, device=dev)
    cluster = torch.tensor([0, 0, 1, 0, 1, 1], device=dev)
    cluster1 = torch.tensor([0, 0, 0, 1], device=dev)
    bad_edge = torch.tensor([[0, 1, 3, 5], [1, 0, 4, 1]], device=dev)        # last edge leaves graph 1
    st = ops.structure_blob(node_ptr, edge_ptr, bad_edge, cluster, 3, 2, c1_ptr, cluster1)
    with pytest.raises(DrgnnError):
        st.sync_counts()
    ok_edge = torch.tensor([[0, 1, 3, 5], [1, 0, 4, 4]], device=dev)
    st = ops.structure_blob(node_ptr, edge_ptr, ok_edge, cluster, 3, 2, c1_ptr, cluster1, out=st)
    st.sync_counts()                                                          # sticky status was re-armed
    short_c1 = torch.tensor([0, 2, 3], dtype=torch.int32, device=dev)        # graph 1 has 2 clusters, 1 id given
    st = ops.structure_blob(node_ptr, edge_ptr, ok_edge, cluster, 3, 2, short_c1, cluster1[:3].contiguous(), out=st)
    with pytest.raises(DrgnnError):
        st.sync_counts()
    assert int(st.blob[48 * 1 + 12 * 3 + 4 * 2 + 5]) == 0                    # blob of graph 1 is marked incomplete


def test_structure_toy_graph_of_reference_test(lib):
    """tests/test_community_pooling.py:10-18,52-58: two copies of the 6-node toy graph with
    clusters [0,0,0,1,1,1 | 2,2,2,3,3,3]: every edge becomes a self loop -> no pooled edge."""
    from deeprank_gnn_b200 import ops
    dev = _dev()
    ei = torch.tensor([[0, 1, 1, 2, 3, 4, 4, 5], [1, 0, 2, 1, 4, 3, 5, 4]])
    edge_index = torch.cat([ei, ei + 6], dim=1).to(dev)
    cluster = torch.tensor([0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3]).to(dev)
    node_ptr = torch.tensor([0, 6, 12], dtype=torch.int32, device=dev)
    edge_ptr = torch.tensor([0, 8, 16], dtype=torch.int32, device=dev)
    st = ops.structure_build(node_ptr, edge_ptr, edge_index, cluster, 6, 8, clusters_are_local=False, mirrors=True)
    K0, E1, K1 = st.sync_counts()
    assert (K0, E1) == (4, 0)
    assert st.cl0.cpu().tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3]
    assert st.batch1[:4].cpu().tolist() == [0, 0, 1, 1]
    x = torch.arange(6, dtype=torch.float32).repeat(2).view(12, 1).to(dev)
    y = torch.empty(4, 1, device=dev)
    arg = torch.empty(4, 1, dtype=torch.int32, device=dev)
    ops.maxpool_fwd(x, st.cmptr0, st.cmem0, y, arg, n_clusters_dev=st.K0_dev)
    assert y.view(-1).cpu().tolist() == [2., 5., 2., 5.]


def test_structure_flags_invalid_input(lib):
    from deeprank_gnn_b200 import ops
    from deeprank_gnn_b200._lib import DrgnnError
    dev = _dev()
    node_ptr = torch.tensor([0, 3, 6], dtype=torch.int32, device=dev)
    edge_ptr = torch.tensor([0, 2, 4], dtype=torch.int32, device=dev)
    edge_index = torch.tensor([[0, 1, 3, 5], [1, 0, 4, 1]], device=dev)       # last edge leaves graph 1
    cluster = torch.tensor([0, 0, 1, 0, 1, 1], device=dev)
    st = ops.structure_build(node_ptr, edge_ptr, edge_index, cluster, 3, 2)
    with pytest.raises(DrgnnError):
        st.sync_counts()


def test_cluster_offset_is_get_preloaded_cluster(lib):
    from deeprank_gnn_b200 import ops
    from oracle import pooling
    g = torch.Generator().manual_seed(0)
    sizes = [5, 1, 17, 300, 2]
    batch = torch.cat([torch.full((n,), i, dtype=torch.long) for i, n in enumerate(sizes)])
    cluster = torch.cat([torch.randint(0, max(1, n // 2 + 1), (n,), generator=g) for n in sizes])
    ref = pooling.get_preloaded_cluster(cluster.clone(), batch)
    ptr = ops.ptr_from_sorted_ids(batch.to(_dev()), len(sizes))
    assert ptr.cpu().tolist() == [0, 5, 6, 23, 323, 325]
    out = ops.cluster_offset_(cluster.clone().to(_dev()), ptr)
    assert torch.equal(out.cpu(), ref)


# ----------------------------------------------------------------------------- aggregation
def _rand_csr(n, max_deg, seed, allow_empty=True):
    g = torch.Generator().manual_seed(seed)
    deg = torch.randint(0 if allow_empty else 1, max_deg + 1, (n,), generator=g)
    rowptr = torch.cat([torch.zeros(1, dtype=torch.long), deg.cumsum(0)])
    col = torch.randint(0, n, (int(rowptr[-1]),), generator=g)
    row = torch.repeat_interleave(torch.arange(n), deg)
    return rowptr.int(), col.int(), row, deg


def _agg_ref(src, row, col, ew, sscale, post_mode, deg):
    n, C = deg.numel(), src.size(1)
    w = torch.ones(col.numel()) if ew is None else ew.clone()
    if sscale is not None:
        w = w * sscale[col.long()]
    out = torch.zeros(n, C).index_add_(0, row, src[col.long()] * w.view(-1, 1))
    if post_mode == 1:
        out = out / deg.clamp(min=1).view(-1, 1).float()
    elif post_mode == 2:
        out = out / deg.view(-1, 1).float()
    return out


@pytest.mark.parametrize('C', [4, 16, 32, 48, 64, 3, 33])
@pytest.mark.parametrize('post_mode', [0, 1, 2])
def test_aggregate_modes(lib, C, post_mode):
    from deeprank_gnn_b200 import ops
    dev = _dev()
    n = 777
    rowptr, col, row, deg = _rand_csr(n, 9, C * 10 + post_mode)
    g = torch.Generator().manual_seed(1)
    src = torch.randn(n, C, generator=g)
    ew = torch.rand(col.numel(), generator=g) + 0.5
    sscale = torch.rand(n, generator=g) + 0.5
    bias = torch.randn(C, generator=g)
    ref = _agg_ref(src, row, col, ew, sscale, post_mode, deg) + bias
    ref = torch.relu(ref)
    out = torch.empty(n, C, device=dev)
    post = torch.empty(n, device=dev)
    ops.aggregate(src.to(dev), rowptr.to(dev), col.to(dev), out, ew=ew.to(dev), sscale=sscale.to(dev),
                  bias=bias.to(dev), post_mode=post_mode, relu=True, post_out=post)
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-5, equal_nan=True)
    if post_mode == 2:
        assert torch.isnan(out.cpu()[deg == 0]).all() or (deg == 0).sum() == 0
    exp_post = {0: torch.ones(n), 1: 1.0 / deg.clamp(min=1).float(), 2: 1.0 / deg.float()}[post_mode]
    torch.testing.assert_close(post.cpu(), exp_post)


def test_aggregate_self_terms_and_slices(lib):
    """sGAT form: left half of a wide buffer = s_i * x_i, right half = weighted neighbour mean."""
    from deeprank_gnn_b200 import ops
    dev = _dev()
    n, C = 500, 16
    rowptr, col, row, deg = _rand_csr(n, 7, 5)
    g = torch.Generator().manual_seed(2)
    x = torch.randn(n, C, generator=g)
    ew = torch.rand(col.numel(), generator=g) + 0.5
    wide = torch.full((n, 2 * C), -7.0, device=dev)
    s = torch.empty(n, device=dev)
    xd = x.to(dev)
    ops.aggregate(xd, rowptr.to(dev), col.to(dev), wide[:, C:], C_=C, ew=ew.to(dev), self_src=xd, self_out=wide[:, :C],
                  selfc_out=s, post_mode=1, self_mode=2)
    d = deg.clamp(min=1).float()
    s_ref = torch.zeros(n).index_add_(0, row, ew) / d
    m_ref = _agg_ref(x, row, col, ew, None, 1, deg)
    torch.testing.assert_close(s.cpu(), s_ref, rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(wide[:, :C].cpu(), x * s_ref.view(-1, 1), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(wide[:, C:].cpu(), m_ref, rtol=1e-5, atol=1e-5)
    # self_mode 3 (backward form) and self_mode 1 added in place
    out = torch.empty(n, C, device=dev)
    ops.aggregate(xd, rowptr.to(dev), col.to(dev), out, self_src=wide[:, :C], selfc_in=s, self_mode=3)
    ref = _agg_ref(x, row, col, None, None, 0, deg) + s_ref.view(-1, 1) * (x * s_ref.view(-1, 1))
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-5, atol=1e-5)
    ops.aggregate(xd, rowptr.to(dev), col.to(dev), out, self_src=xd, self_mode=1)
    torch.testing.assert_close(out.cpu(), _agg_ref(x, row, col, None, None, 0, deg) + x, rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize('C', [16, 32, 64])
def test_aggregate_tiled_equals_rows_kernel(lib, C):
    """Per-graph shared-memory staged kernel == generic kernel, bit for bit (same summation order)."""
    from deeprank_gnn_b200 import ops, synthetic
    from deeprank_gnn_b200.data import Batch
    dev = _dev()
    graphs = synthetic.make_graphs(dict(nodes=(40, 300), edges_per_node=6, feat=C), count=37, seed=9)
    b = Batch.from_data_list(graphs)
    st = ops.structure_build(b._node_ptr.to(dev), b._edge_ptr.to(dev), b.edge_index.to(dev), b.cluster0.to(dev),
                             b._max_n, b._max_e, edge_attr=b.edge_attr.to(dev))
    x = b.x.to(dev)
    o1, o2 = torch.empty_like(x), torch.empty_like(x)
    ops.aggregate(x, st.rowptr0, st.col0, o1, ew=st.w0csr, post_mode=1)
    ops.aggregate(x, st.rowptr0, st.col0, o2, ew=st.w0csr, post_mode=1, tile_ptr=b._node_ptr.to(dev),
                  tile_eptr=b._edge_ptr.to(dev), max_tile_rows=b._max_n, max_tile_edges=b._max_e)
    assert torch.equal(o1, o2)
    o3, o4 = torch.empty_like(x), torch.empty_like(x)
    ops.aggregate(x, st.rowptr0, st.col0, o3)
    ops.aggregate(x, st.rowptr0, st.col0, o4, tile_ptr=b._node_ptr.to(dev), tile_eptr=b._edge_ptr.to(dev),
                  max_tile_rows=b._max_n, max_tile_edges=b._max_e)
    assert torch.equal(o3, o4)
    row, col = b.edge_index
    deg = torch.bincount(row, minlength=x.size(0))
    ref = torch.zeros_like(b.x).index_add_(0, row, b.x[col] * b.edge_attr) / deg.clamp(min=1).view(-1, 1)
    torch.testing.assert_close(o1.cpu(), ref, rtol=1e-5, atol=1e-5)


# ----------------------------------------------------------------------------- dense transform
@pytest.mark.parametrize('math', [0, 1])
@pytest.mark.parametrize('rows,Fin,Fout,groups,layout', [
    (1000, 32, 32, 1, 0), (1000, 16, 32, 2, 0), (333, 3, 32, 1, 0), (64, 64, 128, 1, 0), (64, 128, 1, 1, 0),
    (700, 64, 16, 1, 1), (129, 6, 16, 1, 1), (257, 96, 32, 1, 1), (64, 128, 256, 1, 0), (100, 32, 16, 2, 1),
    (50, 1, 128, 1, 1), (5, 128, 3, 1, 0)])
def test_linear_matches_torch(lib, math, rows, Fin, Fout, groups, layout):
    from deeprank_gnn_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(rows + Fin)
    X = torch.randn(rows, groups * Fin, generator=g)
    W = torch.randn(groups, Fout, Fin, generator=g) / (Fin ** 0.5)       # logical [g][o][k]
    bias = torch.randn(groups * Fout, generator=g)
    mask = (torch.rand(rows, groups * Fout, generator=g) > 0.4).float()
    ref = torch.cat([X[:, i * Fin:(i + 1) * Fin].double() @ W[i].double().t() for i in range(groups)], dim=1)
    ref = torch.relu(ref + bias.double()) * mask.double() * 1.5
    Wst = W if layout == 0 else W.transpose(1, 2)
    out = torch.full((rows + 3, groups * Fout + 5), -1.0, device=dev)     # padded buffer: checks ld handling
    ops.linear(X.to(dev), Wst.contiguous().to(dev), Fin, Fout, out[:rows, :groups * Fout], bias=bias.to(dev), groups=groups,
               w_layout=layout, relu=True, out_mask=mask.to(dev), mask_scale=1.5, math=math)
    torch.testing.assert_close(out[:rows, :groups * Fout].cpu().double(), ref, rtol=2e-5, atol=2e-5)
    assert (out[rows:] == -1).all() and (out[:, groups * Fout:] == -1).all()


@pytest.mark.parametrize('rows,Fin,Fout,layout', [(1000, 32, 32, 0), (128, 32, 16, 1), (129, 64, 64, 0), (70001, 32, 64, 1),
                                                  (5, 8, 16, 0), (40000, 16, 32, 1), (4097, 48, 16, 0)])
def test_linear_tcgen05_matches_torch(lib, rows, Fin, Fout, layout):
    """math = 2: tcgen05.mma kind::tf32 with the 3xTF32 split, fp32 accumulator in TMEM (csrc/linear_tc5.cu) - same
    contract and the same 2e-5 bound as the FMA kernel (bias, ReLU, mask, padded leading dimensions, live rows)."""
    from deeprank_gnn_b200 import _lib, ops
    import ctypes as C
    dev = _dev()
    g = torch.Generator().manual_seed(rows + Fin)
    X = torch.randn(rows, Fin, generator=g)
    W = torch.randn(Fout, Fin, generator=g) / (Fin ** 0.5)
    bias = torch.randn(Fout, generator=g)
    mask = (torch.rand(rows, Fout, generator=g) > 0.4).float()
    ref = torch.relu(X.double() @ W.double().t() + bias.double()) * mask.double() * 1.5
    Wst = W if layout == 0 else W.t()
    out = torch.full((rows + 3, Fout + 4), -1.0, device=dev)              # padded buffer: checks ld handling
    Xd = torch.zeros(rows + 2, Fin + 4, device=dev)
    Xd[:rows, :Fin] = X.to(dev)
    a = _lib.LinearArgs()
    a.X, a.ldx, a.Y, a.ldy, a.groups, a.Fin, a.Fout = Xd.data_ptr(), Fin + 4, out.data_ptr(), Fout + 4, 1, Fin, Fout
    assert _lib.load().drgnn_linear_tcgen05_supported(C.byref(a)) == 1
    ops.linear(Xd[:rows, :Fin], Wst.contiguous().to(dev), Fin, Fout, out[:rows, :Fout], bias=bias.to(dev), w_layout=layout,
               relu=True, out_mask=mask.to(dev), mask_scale=1.5, math=ops.MATH_TCGEN05)
    torch.testing.assert_close(out[:rows, :Fout].cpu().double(), ref, rtol=2e-5, atol=2e-5)
    assert (out[rows:] == -1).all() and (out[:, Fout:] == -1).all()
    # live row count on the device, no bias / activation
    live = torch.tensor([rows // 2 + 1], dtype=torch.int32, device=dev)
    out2 = torch.zeros(rows, Fout, device=dev)
    ops.linear(Xd[:rows, :Fin], Wst.contiguous().to(dev), Fin, Fout, out2, w_layout=layout, rows_dev=live, math=ops.MATH_TCGEN05)
    n = rows // 2 + 1
    torch.testing.assert_close(out2[:n].cpu().double(), (X.double() @ W.double().t())[:n], rtol=2e-5, atol=2e-5)
    assert (out2[n:] == 0).all()


def test_linear_rows_dev_limits_work(lib):
    from deeprank_gnn_b200 import ops
    dev = _dev()
    X = torch.randn(300, 16, device=dev)
    W = torch.randn(8, 16, device=dev)
    out = torch.zeros(300, 8, device=dev)
    live = torch.tensor([130], dtype=torch.int32, device=dev)
    ops.linear(X, W, 16, 8, out, rows_dev=live)
    torch.testing.assert_close(out[:130], X[:130] @ W.t(), rtol=1e-5, atol=1e-5)
    assert (out[130:] == 0).all()


@pytest.mark.parametrize('rows,Fin,Fout,groups,layout', [
    (12800, 32, 32, 1, 0), (3200, 16, 32, 2, 0), (64, 64, 128, 1, 0), (64, 128, 1, 1, 0), (1000, 64, 16, 1, 1),
    (333, 6, 16, 1, 1), (1, 32, 64, 1, 0), (70000, 32, 32, 2, 1), (500, 96, 72, 1, 0)])
def test_linear_wgrad_matches_torch(lib, rows, Fin, Fout, groups, layout):
    from deeprank_gnn_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(rows + Fout)
    X = torch.randn(rows, groups * Fin, generator=g)
    G = torch.randn(rows, groups * Fout, generator=g)
    live = rows - rows // 5
    dW = torch.stack([G[:live, i * Fout:(i + 1) * Fout].double().t() @ X[:live, i * Fin:(i + 1) * Fin].double()
                      for i in range(groups)])                            # [g][o][k]
    db = G[:live].double().sum(0)
    if layout == 1:
        dW = dW.transpose(1, 2)
    out_w = torch.empty(groups * Fin * Fout, device=dev)
    out_b = torch.empty(groups * Fout, device=dev)
    live_dev = torch.tensor([live], dtype=torch.int32, device=dev)
    ops.linear_wgrad(X.to(dev), G.to(dev), Fin, Fout, out_w, out_b, groups=groups, w_layout=layout, rows_dev=live_dev)
    scale = max(1.0, live ** 0.5)
    torch.testing.assert_close(out_w.cpu().double().view(dW.shape) / scale, dW / scale, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(out_b.cpu().double() / scale, db / scale, rtol=1e-4, atol=1e-5)
    # accumulate
    ops.linear_wgrad(X.to(dev), G.to(dev), Fin, Fout, out_w, out_b, groups=groups, w_layout=layout, rows_dev=live_dev,
                     accumulate=True)
    torch.testing.assert_close(out_w.cpu().double().view(dW.shape) / scale, 2 * dW / scale, rtol=1e-4, atol=2e-5)


def test_wgrad_ignores_nan_input_rows_with_zero_gradient(lib):
    from deeprank_gnn_b200 import ops
    dev = _dev()
    X = torch.randn(100, 8, device=dev)
    G = torch.randn(100, 4, device=dev)
    X[7] = float('nan')
    G[7] = 0
    dW = torch.empty(32, device=dev)
    ops.linear_wgrad(X, G, 8, 4, dW)
    Xc = X.clone()
    Xc[7] = 0
    torch.testing.assert_close(dW.view(4, 8), G.t() @ Xc, rtol=1e-5, atol=1e-5)


# ----------------------------------------------------------------------------- pooling / read-out
@pytest.mark.parametrize('C', [16, 32, 64, 5])
def test_maxpool_matches_scatter_max(lib, C):
    from deeprank_gnn_b200 import ops
    dev = _dev()
    g = torch.Generator().manual_seed(C)
    n, K = 1000, 260
    cl = torch.randint(0, K, (n,), generator=g)
    cl[:K] = torch.arange(K)                      # every cluster non-empty
    x = torch.relu(torch.randn(n, C, generator=g))  # ReLU output: many exact ties at 0
    x[5] = float('nan')
    xr = x.clone().requires_grad_(True)
    ref, ref_arg = pyg_min.scatter_max(xr, cl, dim=0, dim_size=K)
    gout = torch.randn(K, C, generator=g)
    ref.backward(gout)
    order = torch.from_numpy(np.argsort(cl.numpy(), kind='stable')).int()
    cmptr = torch.cat([torch.zeros(1, dtype=torch.long), torch.bincount(cl, minlength=K).cumsum(0)]).int()
    y = torch.empty(K, C, device=dev)
    arg = torch.empty(K, C, dtype=torch.int32, device=dev)
    ops.maxpool_fwd(x.to(dev), cmptr.to(dev), order.to(dev), y, arg)
    assert torch.equal(y.cpu(), ref.detach())
    a = arg.cpu().long()
    a[a < 0] = n
    assert torch.equal(a, ref_arg)
    dx = torch.empty(n, C, device=dev)
    ops.maxpool_bwd(gout.to(dev), arg, cl.int().to(dev), dx)
    assert torch.equal(dx.cpu(), xr.grad)
    # fused ReLU gate
    ops.maxpool_bwd(gout.to(dev), arg, cl.int().to(dev), dx, relu_out=x.to(dev))
    assert torch.equal(dx.cpu(), xr.grad * (x > 0))


def test_segment_mean(lib):
    from deeprank_gnn_b200 import ops
    dev = _dev()
    sizes = [3, 0, 17, 1, 40]
    ptr = torch.tensor([0] + list(np.cumsum(sizes)), dtype=torch.int32)
    batch = torch.repeat_interleave(torch.arange(len(sizes)), torch.tensor(sizes))
    x = torch.randn(sum(sizes), 64)
    xr = x.clone().requires_grad_(True)
    ref = pyg_min.scatter_mean(xr, batch, dim=0, dim_size=len(sizes))
    gout = torch.randn(len(sizes), 64)
    ref.backward(gout)
    r = torch.empty(len(sizes), 64, device=dev)
    ops.segment_mean_fwd(x.to(dev), ptr.to(dev), r)
    torch.testing.assert_close(r.cpu(), ref.detach(), rtol=1e-6, atol=1e-6)
    dx = torch.empty_like(x, device=dev)
    ops.segment_mean_bwd(gout.to(dev), ptr.to(dev), dx)
    torch.testing.assert_close(dx.cpu(), xr.grad, rtol=1e-6, atol=1e-7)


# ----------------------------------------------------------------------------- loss / optimiser
def test_mse_and_ce_loss(lib):
    from deeprank_gnn_b200 import ops
    dev = _dev()
    B = 77
    pred = torch.randn(B, requires_grad=True)
    y = torch.rand(B)
    for sig in (False, True):
        pred.grad = None
        p = torch.sigmoid(pred) if sig else pred
        ref = torch.nn.MSELoss()(p, y)
        ref.backward()
        loss = torch.empty(1, device=dev)
        dp = torch.empty(B, device=dev)
        ops.mse_loss(pred.detach().to(dev), y.to(dev), 1.0 / B, loss, dp, sigmoid=sig)
        torch.testing.assert_close(loss.cpu()[0], ref.detach(), rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(dp.cpu(), pred.grad, rtol=1e-5, atol=1e-7)
    logits = torch.randn(B, 3, requires_grad=True)
    tgt = torch.randint(0, 3, (B,))
    w = torch.tensor([0.2, 1.0, 2.5])
    ref = torch.nn.CrossEntropyLoss(weight=w, reduction='mean')(logits, tgt)
    ref.backward()
    loss = torch.empty(1, device=dev)
    dl = torch.empty(B, 3, device=dev)
    ops.ce_loss(logits.detach().to(dev), tgt.to(dev), 1.0 / float(w[tgt].sum()), loss, dl, class_w=w.to(dev))
    torch.testing.assert_close(loss.cpu()[0], ref.detach(), rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dl.cpu(), logits.grad, rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize('n', [10701, 70001])
def test_adam_flat_matches_torch_adam(lib, n):
    from deeprank_gnn_b200 import ops
    dev = _dev()
    p0 = torch.randn(n)
    ref_p = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref_p], lr=0.01)
    p = p0.clone().to(dev)
    m, v = torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    step = torch.zeros(1, device=dev)
    g = torch.Generator().manual_seed(0)
    for _ in range(5):
        grad = torch.randn(n, generator=g)
        grad[::7] = 0
        ref_p.grad = grad.clone()
        opt.step()
        ops.adam_flat(p, grad.to(dev), m, v, step, 0.01)
    torch.testing.assert_close(p.cpu(), ref_p.detach(), rtol=1e-5, atol=1e-6)
    assert float(step.item()) == 5.0


@pytest.mark.parametrize('task', ['mse', 'mse_sigmoid', 'ce', 'none'])
@pytest.mark.parametrize('B,C,Hd,out', [(64, 64, 128, 1), (77, 32, 64, 1), (5, 128, 256, 1), (40, 64, 128, 3), (130, 64, 64, 2)])
def test_fused_head_matches_torch(lib, task, B, C, Hd, out):
    from deeprank_gnn_b200 import ops
    if task == 'ce' and out == 1:
        pytest.skip('cross entropy needs several classes')
    if task in ('mse', 'mse_sigmoid') and out != 1:
        pytest.skip('regression head has one output')
    if not ops.head_fits(C, Hd, out):
        from deeprank_gnn_b200._lib import DrgnnError
        with pytest.raises(DrgnnError):       # too large for one CTA: the engine falls back to the op-by-op path
            ops.head(torch.zeros(B, C, device=_dev()), torch.zeros(Hd, C, device=_dev()), None,
                     torch.zeros(out, Hd, device=_dev()), None, torch.zeros(B, out, device=_dev()))
        return
    dev = _dev()
    g = torch.Generator().manual_seed(B + C)
    R = torch.randn(B, C, generator=g, requires_grad=True)
    fc1, fc2 = torch.nn.Linear(C, Hd), torch.nn.Linear(Hd, out)
    keep = (torch.rand(B, Hd, generator=g) > 0.4).float()
    H = torch.relu(fc1(R)) * keep / 0.6
    pred = fc2(H)
    if task == 'ce':
        tgt = torch.randint(0, out, (B,), generator=g)
        w = torch.rand(out, generator=g) + 0.5
        loss = torch.nn.CrossEntropyLoss(weight=w, reduction='mean')(pred, tgt)
        inv = 1.0 / float(w[tgt].sum())
    elif task == 'none':
        loss = None
    else:
        y = torch.rand(B, generator=g)
        p = torch.sigmoid(pred.reshape(-1)) if task == 'mse_sigmoid' else pred.reshape(-1)
        loss = torch.nn.MSELoss()(p, y)
        inv = 1.0 / B
    if loss is not None:
        loss.backward()
    d = lambda t_: t_.detach().to(dev).contiguous()
    o_pred = torch.empty(B, out, device=dev)
    if task == 'none':
        ops.head(d(R), d(fc1.weight), d(fc1.bias), d(fc2.weight), d(fc2.bias), o_pred, keep=d(keep), keep_scale=1 / 0.6)
        torch.testing.assert_close(o_pred.cpu(), pred.detach(), rtol=1e-4, atol=1e-5)
        return
    o_loss = torch.empty(1, device=dev)
    dW1, db1 = torch.empty(Hd, C, device=dev), torch.empty(Hd, device=dev)
    dW2, db2 = torch.empty(out, Hd, device=dev), torch.empty(out, device=dev)
    dR = torch.empty(B, C, device=dev)
    kw = dict(y_class=tgt.to(dev), class_w=w.to(dev)) if task == 'ce' else dict(y=y.to(dev))
    ops.head(d(R), d(fc1.weight), d(fc1.bias), d(fc2.weight), d(fc2.bias), o_pred,
             task={'mse': ops.TASK_MSE, 'mse_sigmoid': ops.TASK_MSE_SIGMOID, 'ce': ops.TASK_CE}[task], inv_norm=inv,
             keep=d(keep), keep_scale=1 / 0.6, loss=o_loss, dW1=dW1, db1=db1, dW2=dW2, db2=db2, dR=dR, **kw)
    torch.testing.assert_close(o_pred.cpu(), pred.detach(), rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(o_loss.cpu()[0], loss.detach(), rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(dW2.cpu(), fc2.weight.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(db2.cpu(), fc2.bias.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(dW1.cpu(), fc1.weight.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(db1.cpu(), fc1.bias.grad, rtol=1e-4, atol=1e-6)
    torch.testing.assert_close(dR.cpu(), R.grad, rtol=1e-4, atol=1e-6)
