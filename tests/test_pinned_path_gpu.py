"""Parity of the paths bench.py actually times ("pin the benchmarked path"):

  * ``Engine.train_resident`` (chunk CUDA graphs, side-stream structure passes, slot recycling, tail chunk)
    against the same steps issued one by one;
  * the cluster whole-step kernel + in-kernel reduction at the benchmarked batch size (B = 64) through
    ``PackedBatch(idx16)`` against the CPU oracle - predictions to ABSOLUTE 1e-4 (north_star), gradients to
    1e-4 of their scale;
  * cfg3 / cfg4 / cfg5 at their per-GPU batch sizes (64 / 32 / 64) against the vectorised oracle;
  * a malformed batch in the middle of a pass is reported by ``validate()`` and does not poison later steps;
  * SURVEY 8a-ter leftovers: ``number_edge_features > 1`` (GINet), 1-D ``edge_attr``, a ``Batch`` without
    ``cluster0`` through ``Engine``.
"""
import copy

import pytest
import torch

from test_engine_gpu import _close, _device_batch, _engine, _oracle_run

pytestmark = pytest.mark.gpu


def _close_abs(a, b, name, tol=1e-4):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, '%s: shape %s vs %s' % (name, tuple(a.shape), tuple(b.shape))
    err = float((a - b).abs().max())
    assert err <= tol, '%s: max|diff| = %.3e (absolute bound %.1e)' % (name, err, tol)


def _pool(cfg, n_batches, B, seed, edge_attr=False, compact=None):
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.data import Batch, PackedBatch
    graphs = synthetic.make_graphs(cfg, count=2 * B, seed=seed, internal=False)
    g = torch.Generator().manual_seed(seed)
    packed = []
    for _ in range(n_batches):
        idx = torch.randperm(len(graphs), generator=g)[:B].tolist()
        packed.append(PackedBatch.from_batch(Batch.from_data_list([graphs[i] for i in idx]), idx16=True,
                                             edge_attr=edge_attr, compact=compact))
    return packed


@pytest.mark.parametrize('net', ['GINet', 'sGAT'])
def test_train_resident_with_the_first_aggregation_in_the_structure_pass(lib, net):
    """The resident loop with the first aggregation computed by the structure pass (``pre_agg``) leaves exactly the
    weights, optimiser state and loss of the loop whose step kernels aggregate themselves."""
    from deeprank_gnn_b200.engine import Engine
    packed = _pool('cfg2' if net == 'GINet' else 'cfg3', 16, 64, seed=5, edge_attr=net == 'sGAT')
    ref = None
    for pre in ('0', '1'):
        e = Engine(net, 32, 1, 1, device='cuda:0', seed=3, lr=1e-3, graph=True)
        e.pre_agg = pre
        r = [e.upload(pb, slot=i) for i, pb in enumerate(packed)]
        loss, pred = e.train_resident(r, steps=37)
        e.validate()
        assert (e.structs[r[0].sslot].zin1 is not None) == (pre == '1')
        got = (loss.clone(), pred.clone(), e.params.data.clone(), e.exp_avg.clone(), e.exp_avg_sq.clone())
        if ref is None:
            ref = got
        else:
            for a_, b_ in zip(got, ref):
                assert torch.equal(a_, b_), pre


@pytest.mark.parametrize('steps', [40, 43, 5])
def test_train_resident_chunk_graphs_equal_single_steps(lib, steps):
    """bench.py's `value` path: 16 resident batches of 64 graphs, ``train_resident`` (chunk graphs of 16 steps +
    a tail chunk, structure passes two steps ahead on side streams, 4 structure slots recycled) must leave
    exactly the weights, optimiser state and loss the same steps leave when issued one at a time."""
    from deeprank_gnn_b200.engine import Engine
    packed = _pool('cfg2', 16, 64, seed=7)
    ea = Engine('GINet', 32, 1, 1, device='cuda:0', seed=3, lr=1e-3, graph=True)
    eb = Engine('GINet', 32, 1, 1, device='cuda:0', seed=3, lr=1e-3, graph=True)
    ra = [ea.upload(pb, slot=i) for i, pb in enumerate(packed)]
    rb = [eb.upload(pb, slot=i) for i, pb in enumerate(packed)]
    la, pa = ea.train_resident(ra, steps=steps)
    assert ea._rotation_chunk(ra) == 16
    for i in range(steps):
        lb, pb_ = eb.step(rb[i % len(rb)])
    ea.validate(), eb.validate()
    assert float(ea.step_dev[0]) == float(steps) == float(eb.step_dev[0])
    assert torch.equal(la, lb) and torch.equal(pa, pb_)
    assert torch.equal(ea.params.data, eb.params.data)
    assert torch.equal(ea.exp_avg, eb.exp_avg) and torch.equal(ea.exp_avg_sq, eb.exp_avg_sq)
    # a second call that continues the rotation (bench.py: align chunk, then the timed chunk) stays equal
    la, pa = ea.train_resident(ra, steps=20, start=steps)
    for i in range(steps, steps + 20):
        lb, pb_ = eb.step(rb[i % len(rb)])
    assert torch.equal(la, lb) and torch.equal(ea.params.data, eb.params.data)


def test_ginet_cfg2_batch64_packed_cluster_kernel_matches_oracle(lib):
    """The benchmarked configuration itself: B = 64 (128 CTAs, grid barrier, in-kernel 4-quarter reduction +
    Adam) through PackedBatch(idx16) and the CUDA-graph path, against the CPU oracle."""
    from deeprank_gnn_b200 import _lib, ops, synthetic
    from deeprank_gnn_b200.data import Batch, PackedBatch
    graphs = synthetic.make_graphs('cfg2', count=64, seed=11, internal=False)
    sd0, loss, pred, grads, sd1 = _oracle_run('GINet', graphs, (16, 32), 1, train=False, lr=1e-3)
    eng = _engine('GINet', graphs, (16, 32), 1, sd0, lr=1e-3, graph=True).eval()
    pb = PackedBatch.from_batch(Batch.from_data_list(graphs), idx16=True, edge_attr=False)
    d = eng.upload(pb)
    eloss, epred = eng.step(d)
    eng.validate()
    assert ops.ginet_step_last_variant() == 2 and _lib.load().drgnn_ginet_step_last_launches() == 1
    assert eng.structs[d.sslot].blob_only
    _close_abs(epred.view(-1), pred, 'pred')
    _close_abs(eloss.view(-1), loss.view(-1), 'loss')
    for name, g in eng.named_grads().items():
        _close(g, grads[name], 'grad ' + name)
    for name, p in eng.state_dict().items():
        solid = grads[name].abs() > 1e-5
        if solid.any():
            _close_abs(p.cpu()[solid], sd1[name][solid], 'param ' + name)


@pytest.mark.parametrize('net,cfg,B', [('sGAT', 'cfg3', 64), ('GINet', 'cfg4', 32), ('FoutNet', 'cfg5', 64)])
def test_other_configs_at_their_per_gpu_batch_match_oracle(lib, net, cfg, B):
    """BASELINE configs 3-5 at the batch one GPU sees (64 / 256 over 8 / 512 over 8) against the loop-free
    oracle (proven equal to the literal loops in tests/test_oracle.py)."""
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.data import Batch, PackedBatch
    c = synthetic.CONFIGS[cfg]
    graphs = synthetic.make_graphs(cfg, count=B, seed=17, internal=False)
    sd0, loss, pred, grads, sd1 = _oracle_run(net, graphs, c['hidden'], 1, train=False, literal=False)
    eng = _engine(net, graphs, c['hidden'], 1, sd0, graph=True).eval()
    pb = PackedBatch.from_batch(Batch.from_data_list(graphs), idx16=True, edge_attr=net == 'sGAT')
    eloss, epred = eng.step(eng.upload(pb))
    eng.validate()
    _close_abs(epred.view(-1), pred, 'pred')
    _close_abs(eloss.view(-1), loss.view(-1), 'loss')
    for name, g in eng.named_grads().items():
        _close(g, grads[name], 'grad ' + name)


@pytest.mark.parametrize('fault', ['edge_outside', 'cluster_range'])
def test_malformed_batch_in_the_middle_of_a_pass_is_reported_and_contained(lib, fault):
    """(a) a bad batch that is not the last one of a pass must still be reported by validate();
    (b) an invalid graph must reach the grid barrier of the fused reduction (else the barrier counters stay
    armed wrong and every later step reduces garbage): after the error a clean step equals a fresh engine's."""
    from deeprank_gnn_b200._lib import DrgnnError
    from deeprank_gnn_b200.engine import Engine
    # (uint16 cluster ids of the compact records cannot leave the bitmap range: that fault needs int32 ids)
    packed = _pool('cfg2', 9, 12, seed=23, compact=None if fault == 'edge_outside' else False)
    bad = packed[4]
    v = bad.views(bad.buf)
    if fault == 'edge_outside':
        v['edge_index'][0, 5] = 30000                      # uint16 local id far beyond the graph
    else:
        v['cluster0'][3] = 1 << 20                          # id range beyond the bitmap cap: blob left incomplete
    eng = Engine('GINet', 32, 1, 1, device='cuda:0', seed=5, lr=1e-3, graph=True)
    eng.train_batches(packed)
    with pytest.raises(DrgnnError):
        eng.validate()
    eng.validate()                                          # reported once, then re-armed
    fresh = Engine('GINet', 32, 1, 1, device='cuda:0', seed=5, lr=1e-3, graph=True)
    fresh.load_state_dict(eng.state_dict())
    fresh.exp_avg.copy_(eng.exp_avg), fresh.exp_avg_sq.copy_(eng.exp_avg_sq)
    fresh.step_dev[0] = eng.step_dev[0]
    l1, p1 = eng.step(eng.upload(packed[0]))
    l2, p2 = fresh.step(fresh.upload(packed[0]))
    eng.validate(), fresh.validate()
    assert torch.equal(l1, l2) and torch.equal(p1, p2)
    assert torch.equal(eng.grads, fresh.grads) and torch.equal(eng.params.data, fresh.params.data)


def test_ginet_with_two_edge_features_matches_oracle(lib):
    """number_edge_features = 2 (ginet.py:24-28): the attention path is still the identity (softmax over one
    column), only the dead parameters change shape: fc_edge_attr [2,2], fc_attention [1, 2*out+2]."""
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.engine import Engine
    from helpers import to_oracle_batch
    from oracle import nets as onets
    from oracle import step as ostep
    graphs = synthetic.make_graphs('cfg2', count=6, seed=3, internal=False)
    for g in graphs:
        g.edge_attr = torch.cat([g.edge_attr, g.edge_attr * 0.5 + 0.1], dim=1).contiguous()
    torch.manual_seed(1)
    model = onets.GINet(32, 1, 2).eval()
    sd0 = copy.deepcopy(model.state_dict())
    assert tuple(sd0['conv1.fc_edge_attr.weight'].shape) == (2, 2)
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    loss, pred = ostep.train_step(model, opt, ostep.make_loss('reg'), to_oracle_batch(graphs))
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    eng = Engine('GINet', 32, 1, 2, device='cuda:0').eval()
    eng.load_state_dict(sd0)
    eloss, epred = eng.step(_device_batch(graphs))
    eng.validate()
    _close_abs(epred.view(-1), pred, 'pred')
    _close_abs(eloss.view(-1), loss.view(-1), 'loss')
    for name, g in eng.named_grads().items():
        _close(g, grads[name], 'grad ' + name)


@pytest.mark.parametrize('net', ['GINet', 'sGAT'])
def test_one_dimensional_edge_attr_is_unsqueezed(lib, net):
    """ginet.py:54-55 / sGAT.py:66-67: a 1-D edge_attr is treated as [E, 1]."""
    from deeprank_gnn_b200 import synthetic
    graphs = synthetic.make_graphs('cfg2', count=5, seed=9, internal=False)
    sd0, loss, pred, grads, sd1 = _oracle_run(net, graphs, (16, 32), 1, train=False)
    flat = copy.deepcopy(graphs)
    for g in flat:
        g.edge_attr = g.edge_attr.view(-1).contiguous()
    eng = _engine(net, graphs, (16, 32), 1, sd0).eval()
    eloss, epred = eng.step(_device_batch(flat))
    eng.validate()
    _close_abs(epred.view(-1), pred, 'pred')
    for name, g in eng.named_grads().items():
        _close(g, grads[name], 'grad ' + name)


def test_batch_without_clusters_is_refused_like_the_reference(lib):
    """The reference networks read data.cluster0 / cluster1 (ginet.py:106-107): a Batch without them fails there
    with an AttributeError; the engine refuses it with a message that names the missing step."""
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200._lib import DrgnnError
    from deeprank_gnn_b200.data import Batch
    from deeprank_gnn_b200.engine import DeviceBatch
    graphs = synthetic.make_graphs('cfg2', count=3, seed=2, internal=False)
    for g in graphs:
        del g.cluster0, g.cluster1
    with pytest.raises(DrgnnError, match='cluster0'):
        DeviceBatch.from_batch(Batch.from_data_list(graphs), 'cuda:0')


@pytest.mark.parametrize('case', ['GINet-blob', 'sGAT-blob', 'FoutNet-large', 'FoutNet-large-full'])
def test_compact_records_give_the_same_structure_and_steps(lib, case):
    """Compact feeder records (first half of every graph's mirrored edge list, uint16 cluster ids) against the
    full uint16 records: the structure pass rebuilds bit-identical blobs (bitmap pass and counting-sort pass) and
    training steps leave bit-identical predictions, loss and weights."""
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.data import Batch, PackedBatch
    from deeprank_gnn_b200.engine import Engine
    net = case.split('-')[0]
    cfg = dict(nodes=(300, 600), edges_per_node=8, feat=32) if 'large' in case else dict(nodes=(20, 200), edges_per_node=5, feat=32)
    graphs = synthetic.make_graphs(cfg, count=12, seed=17, internal=False)
    b = Batch.from_data_list(graphs)
    full = PackedBatch.from_batch(b, idx16=True, edge_attr=net == 'sGAT', compact=False)
    half = PackedBatch.from_batch(b, idx16=True, edge_attr=net == 'sGAT')
    assert half.compact and not full.compact and half.nbytes < full.nbytes
    ea = Engine(net, 32, 1, 1, device='cuda:0', seed=3, lr=1e-3, dropout=0.0)
    eb = Engine(net, 32, 1, 1, device='cuda:0', seed=3, lr=1e-3, dropout=0.0)
    if case.endswith('full'):             # the counting-sort pass (graph_local_kernel + finalize) instead of the blob pass
        ea.blob_structure = eb.blob_structure = False
    da, db = ea.upload(full), eb.upload(half)
    assert db.edge_half and not da.edge_half and db.cluster0.dtype == torch.int16
    for step in range(3):
        la, pa = ea.step(da)
        lb, pb_ = eb.step(db)
        ea.validate(), eb.validate()
        sa, sb = ea.structs[da.sslot], eb.structs[db.sslot]
        assert sa.blob_only == sb.blob_only == (not case.endswith('full'))
        assert torch.equal(sa.blob, sb.blob)
        assert torch.equal(torch.nan_to_num(pa, nan=-7.0), torch.nan_to_num(pb_, nan=-7.0))
        assert torch.equal(torch.nan_to_num(la, nan=-7.0), torch.nan_to_num(lb, nan=-7.0))
    assert torch.equal(torch.nan_to_num(ea.params.data, nan=-7.0), torch.nan_to_num(eb.params.data, nan=-7.0))
