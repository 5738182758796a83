"""Property tests (hypothesis) of the host-side record formats and the sharding rule, on randomly drawn ragged
batches: whatever the sizes, every packed form decodes to the batch it was made from, and every shard plan covers
the graphs exactly once.  CPU only."""
import os

import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from conftest import ROOT  # noqa: F401

SET = dict(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])


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


@settings(max_examples=10, deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture])
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
