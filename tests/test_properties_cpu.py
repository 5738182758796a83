"""Property tests (hypothesis) of the host-side record formats and the sharding rule, on randomly drawn ragged
batches: whatever the sizes, every packed form decodes to the batch it was made from, and every shard plan covers
the graphs exactly once.  CPU only."""
import os

import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import ROOT  # noqa: F401

# derandomize: the same examples on every run (the suite must not flake in the driver's -x run); database off
SET = dict(max_examples=25, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow])


def _batch(lo, hi, count, seed, feat=8):
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.data import Batch
    graphs = synthetic.make_graphs(dict(nodes=(lo, hi), edges_per_node=4, feat=feat), count=count, seed=seed)
    return graphs, Batch.from_data_list(graphs)


def _decode_edges(pb, v, batch):
    """edge_index of a packed record as global int64 ids (the inverse of the three encodings)."""
    ei = v['edge_index'].long()
    if not pb.idx16:
        return ei
    np_, ep_ = batch._node_ptr.long(), batch._edge_ptr.long()
    if not pb.compact:
        first = torch.repeat_interleave(np_[:-1], ep_[1:] - ep_[:-1])
        return ei + first
    out = torch.empty(2, int(ep_[-1]), dtype=torch.int64)
    for g in range(batch.num_graphs):
        e0, e1 = int(ep_[g]), int(ep_[g + 1])
        h = (e1 - e0) // 2
        half = ei[:, e0 // 2:e0 // 2 + h] + int(np_[g])
        out[:, e0:e0 + h] = half
        out[:, e0 + h:e1] = half.flip(0)
    return out


@settings(**SET)
@given(lo=st.integers(4, 40), span=st.integers(0, 120), count=st.integers(1, 9), seed=st.integers(0, 10 ** 6),
       mode=st.sampled_from(['i32', 'idx16', 'compact']), attrs=st.booleans())
def test_every_packed_form_decodes_to_its_batch(lo, span, count, seed, mode, attrs):
    from deeprank_gnn_b200.data import PackedBatch
    _graphs, b = _batch(lo, lo + span, count, seed)
    kw = dict(pin=False, edge_attr=attrs)
    if mode == 'idx16':
        kw.update(idx16=True, compact=False)
    elif mode == 'compact':
        kw.update(idx16=True)
    pb = PackedBatch.from_batch(b, **kw)
    v = pb.views(pb.buf)
    assert pb.nbytes == 4 * pb.numel and pb.numel <= pb.capacity_numel
    assert torch.equal(v['x'], b.x) and torch.equal(v['y'], b.y)
    if attrs:
        assert torch.equal(v['edge_attr'].reshape(b.edge_attr.shape), b.edge_attr)
    else:
        assert v['edge_attr'] is None
    assert torch.equal(v['node_ptr'], b._node_ptr) and torch.equal(v['edge_ptr'], b._edge_ptr)
    assert torch.equal(v['cluster0'].long(), b.cluster0) and torch.equal(v['cluster1'].long(), b.cluster1)
    assert torch.equal(_decode_edges(pb, v, b), b.edge_index)
    # a staging buffer of capacity size (what Engine.upload allocates) exposes the same sections at the same offsets
    dev_like = torch.zeros(pb.capacity_numel)
    dev_like[:pb.numel] = pb.buf
    w = pb.views(dev_like, capacity=True)
    assert torch.equal(w['x'], b.x) and torch.equal(w['cluster1'][:b.cluster1.numel()].long(), b.cluster1)
    assert w['cluster1'].numel() == b.x.size(0)
    # batches of the same shape share one layout (one CUDA graph serves both); a different shape does not
    _g2, b2 = _batch(lo, lo + span, count, seed + 1)
    same_shape = (b2.x.size(0), b2.edge_index.size(1), b2._max_n, b2._max_e) == \
        (b.x.size(0), b.edge_index.size(1), b._max_n, b._max_e)
    pb2 = PackedBatch.from_batch(b2, **kw)
    if same_shape and pb2.compact == pb.compact:
        for k in PackedBatch.FLOAT_SECTIONS:
            assert pb2.offsets[k][0] == pb.offsets[k][0]


@settings(max_examples=10, deadline=None, derandomize=True, database=None,
          suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
@given(seed=st.integers(0, 10 ** 6), n=st.integers(1, 5))
def test_packed_cache_returns_the_records_it_was_built_from(tmp_path_factory, seed, n):
    from deeprank_gnn_b200.data import PackedBatch, PackedCache
    recs = []
    for i in range(n):
        _g, b = _batch(5, 60, 1 + (seed + i) % 4, seed + i)
        recs.append(PackedBatch.from_batch(b, pin=False, idx16=True, edge_attr=bool(i % 2)))
    path = str(tmp_path_factory.mktemp('cache') / 'records.pack')
    PackedCache.build(path, recs)
    cache = PackedCache(path)
    try:
        assert len(cache) == n
        for got, want in zip(cache, recs):
            assert got.layout_key() == want.layout_key() and got.compact == want.compact
            assert torch.equal(got.buf[:want.numel], want.buf[:want.numel])
    finally:
        cache.close()


@settings(**SET)
@given(costs=st.lists(st.integers(1, 5000), min_size=1, max_size=200), world=st.sampled_from([1, 2, 4, 8]),
       balance=st.booleans())
def test_shard_plan_covers_every_graph_once(costs, world, balance):
    """parallel.shard_indices: equal counts where the batch divides, every graph in exactly one shard,
    and with balancing no shard heavier than the LPT bound (mean + largest item)."""
    from deeprank_gnn_b200.parallel import shard_indices
    parts = shard_indices(costs, world, balance)
    assert len(parts) == world
    flat = sorted(i for p in parts for i in p)
    assert flat == list(range(len(costs)))
    sizes = [len(p) for p in parts]
    assert max(sizes) - min(sizes) <= 1
    if balance and len(costs) >= world:
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) <= sum(costs) / world + max(costs) * 2


# ----------------------------------------------------------------------------------------------- oracle invariants
def _oracle_batch(graphs):
    from oracle import pyg_min
    return pyg_min.Batch.from_data_list(
        [pyg_min.Data(**{k: (g[k].clone() if torch.is_tensor(g[k]) else g[k]) for k in g.keys}) for g in graphs])


@settings(**SET)
@given(lo=st.integers(4, 30), span=st.integers(0, 80), count=st.integers(1, 6), seed=st.integers(0, 10 ** 6))
def test_oracle_pooling_invariants(lo, span, count, seed):
    """Size-independent properties of get_preloaded_cluster + community_pooling (community_pooling.py:25-30, 161-251)
    the CUDA structure pass is held to bit for bit: dense relabelling is idempotent and order preserving, the pooled
    edge list is sorted by (row, col), unique, free of self-loops, symmetric for symmetric input, its attributes
    conserve the attribute mass of the edges that survive, the pooled batch vector is sorted, and every pooled
    feature is the maximum over its cluster."""
    from oracle import pooling, pyg_min
    graphs, _b = _batch(lo, lo + span, count, seed, feat=4)
    ob = _oracle_batch(graphs)
    cl = pooling.get_preloaded_cluster(ob.cluster0.clone(), ob.batch)
    assert torch.equal(cl, pooling.get_preloaded_cluster_closed_form(ob.cluster0.clone(), ob.batch))
    dense, perm = pyg_min.consecutive_cluster(cl)
    again, _ = pyg_min.consecutive_cluster(dense)
    assert torch.equal(again, dense)                                         # idempotent
    assert torch.equal(torch.unique(dense), torch.arange(int(dense.max()) + 1))   # no gaps
    order = torch.argsort(cl, stable=True)
    assert bool((dense[order][1:] >= dense[order][:-1]).all())               # order preserving
    x0, ei0, ea0 = ob.x.clone(), ob.edge_index.clone(), ob.edge_attr.clone()
    out = pooling.community_pooling(cl, ob)
    K = int(dense.max()) + 1
    assert out.x.shape == (K, x0.size(1)) and out.batch.numel() == K
    assert bool((out.batch[1:] >= out.batch[:-1]).all())
    for k in range(0, K, max(1, K // 7)):                                    # a few clusters: the maximum of the members
        assert torch.equal(out.x[k], x0[dense == k].max(dim=0).values)
    pe = out.edge_index
    if pe.numel():
        key = pe[0] * K + pe[1]
        assert bool((key[1:] > key[:-1]).all())                              # sorted by (row, col) and unique
        assert bool((pe[0] != pe[1]).all())                                  # no self-loops
        back = pe[1] * K + pe[0]
        assert torch.equal(torch.sort(back).values, key)                     # symmetric input stays symmetric
    keep = dense[ei0[0]] != dense[ei0[1]]
    assert pe.size(1) == torch.unique(dense[ei0[0]][keep] * K + dense[ei0[1]][keep]).numel()
    torch.testing.assert_close(out.edge_attr.sum(), ea0[keep].sum(), rtol=1e-5, atol=1e-5)   # coalesce sums attributes


@settings(**SET)
@given(n=st.integers(1, 300), c=st.integers(1, 9), k=st.integers(1, 40), seed=st.integers(0, 10 ** 6))
def test_oracle_scatter_reductions_against_a_loop(n, c, k, seed):
    """scatter_sum / scatter_mean / scatter_max (torch_scatter's published semantics) against the
    per-segment loop: sums in index order, mean over max(count, 1), max with untouched rows = 0."""
    from oracle import pyg_min
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(n, c, generator=g)
    idx = torch.randint(0, k, (n,), generator=g)
    s = pyg_min.scatter_sum(src, idx, dim=0, dim_size=k)
    m = pyg_min.scatter_mean(src, idx, dim=0, dim_size=k)
    mx, arg = pyg_min.scatter_max(src, idx, dim=0, dim_size=k)
    for j in range(k):
        rows = src[idx == j]
        torch.testing.assert_close(s[j], rows.sum(0) if rows.numel() else torch.zeros(c), rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(m[j], rows.mean(0) if rows.numel() else torch.zeros(c), rtol=1e-5, atol=1e-5)
        assert torch.equal(mx[j], rows.max(0).values if rows.numel() else torch.zeros(c))
        if rows.numel():
            assert torch.equal(src[arg[j], torch.arange(c)], mx[j])
