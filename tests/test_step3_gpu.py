"""The general cluster whole-step kernel (csrc/fused_step3.cuh, ``drgnn_net_step``): GINet / sGAT / FoutNet,
one graph per thread-block cluster, nodes tiled over the cluster's CTAs (neighbours through DSMEM).

  * against the op-level path (aggregate / linear / maxpool / ... launches) on the same batch: every
    intermediate (zin1, Z1, argmax0, zin2, Z2, argmax1, read-out), predictions, loss, every gradient, for
    1, 2 and 4 node tiles - this is what localises a wrong phase;
  * against the CPU oracle at the BASELINE shapes (cfg3 sGAT, cfg4 GINet 500 nodes hidden 32/64, cfg5 FoutNet
    50-1000 nodes), predictions to absolute 1e-4;
  * the Fout NaN rule, degenerate graphs, classification + injected dropout, scoring;
  * in-kernel gradient reduction + Adam vs the reduction launch; CUDA-graph replay (train_resident).
"""
import pytest
import torch

from test_engine_gpu import _close, _device_batch, _engine, _oracle_run
from test_pinned_path_gpu import _close_abs

pytestmark = pytest.mark.gpu

HIDDEN = (16, 32)


def _pair(net, hidden=HIDDEN, out=1, tiles=0, **kw):
    """(op-level engine, step3 engine) with identical weights."""
    from deeprank_gnn_b200.engine import Engine
    e_o = Engine(net, 32, out, 1, hidden=hidden, device='cuda:0', seed=5, dropout=0.0, fused_graph=False, fused_head=False,
                 **kw)
    e_o.step3 = False
    e_f = Engine(net, 32, out, 1, hidden=hidden, device='cuda:0', seed=5, dropout=0.0, **kw)
    e_f.step3_tiles = tiles
    return e_o, e_f


@pytest.mark.parametrize('tiles', [1, 2, 4])
@pytest.mark.parametrize('net', ['GINet', 'sGAT', 'FoutNet'])
def test_step3_intermediates_and_gradients_equal_op_level_path(lib, net, tiles):
    from deeprank_gnn_b200 import ops, synthetic
    graphs = synthetic.make_graphs(dict(nodes=(5, 200), edges_per_node=5, feat=32), count=21, seed=13)
    d = _device_batch(graphs)
    e_o, e_f = _pair(net, tiles=tiles)
    e_f.keep_intermediates = True
    e_f.fused_tc = tiles != 1             # 1 tile: fp32 FMA tiles (bit-identical to the op-level kernels for GINet);
                                          # 2 / 4 tiles: the tensor-core (3xTF32) products
    for step in range(3):
        lo, po = e_o.step(d)
        lf, pf = e_f.step(d)
        e_f.validate(), e_o.validate()
        assert e_f._last_path == 'step3' and e_o._last_path == 'ops'
        assert ops.net_step_last()[1] == tiles
        if step == 0:
            N, (K0, _E1, K1) = d.N, e_o.structs[0].sync_counts()
            exact = net == 'GINet' and tiles == 1          # same fmaf chains, same summation orders
            for name, rows in (('Zin1', N), ('Z1', N), ('Zin2', K0), ('Z2', K0)):
                a, b = getattr(e_f.ws, name)[:rows], getattr(e_o.ws, name)[:rows]
                if exact:
                    assert torch.equal(a, b), name
                else:
                    tol = dict(rtol=3e-5, atol=1e-5) if e_f.fused_tc else dict(rtol=1e-5, atol=1e-6)
                    torch.testing.assert_close(a, b, equal_nan=True, msg=name, **tol)
            for name, rows in (('arg0', K0), ('arg1', K1)):
                a, b = getattr(e_f.ws, name)[:rows], getattr(e_o.ws, name)[:rows]
                same = float((a == b).float().mean())
                assert same == 1.0 if exact else same > 0.995, '%s: %.5f equal' % (name, same)
            torch.testing.assert_close(e_f.ws.R[:d.B], e_o.ws.R[:d.B], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(pf, po, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(lf, lo, rtol=1e-4, atol=1e-6)
        gf, go = e_f.named_grads(), e_o.named_grads()
        for name in gf:
            torch.testing.assert_close(gf[name], go[name], rtol=1e-3, atol=1e-5, msg='%s step %d' % (name, step))
    torch.testing.assert_close(e_f.params.data, e_o.params.data, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize('net,cfg,count', [('sGAT', 'cfg3', 16), ('GINet', 'cfg4', 6), ('FoutNet', 'cfg5', 8),
                                           ('FoutNet', 'cfg3', 16), ('sGAT', 'cfg5', 6)])
def test_step3_matches_oracle_at_baseline_shapes(lib, net, cfg, count):
    from deeprank_gnn_b200 import ops, synthetic
    from deeprank_gnn_b200.data import Batch, PackedBatch
    c = synthetic.CONFIGS[cfg]
    graphs = synthetic.make_graphs(cfg, count=count, seed=4, internal=False)
    sd0, loss, pred, grads, sd1 = _oracle_run(net, graphs, c['hidden'], 1, train=False, literal=False)
    eng = _engine(net, graphs, c['hidden'], 1, sd0, graph=True).eval()
    pb = PackedBatch.from_batch(Batch.from_data_list(graphs), idx16=True, edge_attr=net == 'sGAT')
    eloss, epred = eng.step(eng.upload(pb))
    eng.validate()
    assert eng._last_path == 'step3', eng._last_path
    _close_abs(epred.view(-1), pred, 'pred')
    _close_abs(eloss.view(-1), loss.view(-1), 'loss')
    for name, g in eng.named_grads().items():
        _close(g, grads[name], 'grad ' + name)
    if cfg in ('cfg4', 'cfg5'):
        assert ops.net_step_last()[1] > 1            # these graphs need several CTAs per graph


def test_step3_fout_isolated_node_and_degenerate_graphs(lib):
    """Node without neighbour (Fout: NaN row that no maximum selects), a graph swallowed by one cluster (empty
    pooled graph: conv2 sees no neighbour -> Fout NaN rows at level 1 -> zero read-out), singleton clusters with
    id gaps, B = 1."""
    from deeprank_gnn_b200 import synthetic
    graphs = synthetic.make_graphs(dict(nodes=(40, 120), edges_per_node=5, feat=32), count=4, seed=41, internal=False)
    g0 = graphs[0]
    keep = (g0.edge_index[0] != 7) & (g0.edge_index[1] != 7)
    g0.edge_index, g0.edge_attr = g0.edge_index[:, keep].contiguous(), g0.edge_attr[keep].contiguous()
    g1 = graphs[1]
    g1.cluster0 = torch.zeros(g1.x.size(0), dtype=torch.long)
    g1.cluster1 = torch.zeros(1, dtype=torch.long)
    g2 = graphs[2]
    n2 = g2.x.size(0)
    g2.cluster0 = torch.arange(n2, dtype=torch.long) * 3 + 7
    g2.cluster1 = (torch.arange(n2, dtype=torch.long) // 2) * 5
    for net in ('FoutNet', 'sGAT', 'GINet'):
        for subset in (graphs, graphs[1:2], graphs[:1]):
            sd0, loss, pred, grads, sd1 = _oracle_run(net, subset, HIDDEN, 1, train=False)
            for tiles in (1, 2):
                eng = _engine(net, subset, HIDDEN, 1, sd0).eval()
                eng.step3_tiles = tiles
                eloss, epred = eng.step(_device_batch(subset))
                eng.validate()
                assert eng._last_path == 'step3'
                _close_abs(epred.view(-1), pred, '%s pred' % net)
                for name, g in eng.named_grads().items():
                    assert torch.isfinite(g).all(), name
                    _close(g, grads[name], '%s grad %s' % (net, name))


def test_step3_classification_dropout_and_scoring(lib):
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.engine import Engine
    graphs = synthetic.make_graphs('cfg2', count=12, seed=21, internal=False)
    for i, g in enumerate(graphs):
        g.y = torch.tensor([float(i % 3)])
    w = torch.tensor([0.5, 1.5, 1.0])
    d = _device_batch(graphs, classes=[0, 1, 2])
    inv = 1.0 / float(w[d.y_class.cpu()].sum())
    keep = (torch.rand(12, 64, generator=torch.Generator().manual_seed(1)) > 0.3).float()
    kw = dict(device='cuda:0', seed=9, task='class', class_weights=w, dropout=0.3)
    e_o = Engine('sGAT', 32, 3, 1, fused_graph=False, fused_head=False, **kw)
    e_o.step3 = False
    e_f = Engine('sGAT', 32, 3, 1, **kw)
    lo, po = e_o.step(d, inv_norm=inv, keep_mask=keep)
    lf, pf = e_f.step(d, inv_norm=inv, keep_mask=keep)
    assert e_f._last_path == 'step3' and e_o._last_path == 'ops'
    torch.testing.assert_close(pf, po, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(lf, lo, rtol=1e-4, atol=1e-6)
    for name, gfv in e_f.named_grads().items():
        torch.testing.assert_close(gfv, e_o.named_grads()[name], rtol=1e-3, atol=1e-5, msg=name)
    pe, po2 = e_f.eval().forward(d), e_o.eval().forward(d)
    torch.testing.assert_close(pe, po2, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('net', ['sGAT', 'FoutNet'])
def test_step3_in_kernel_reduction_equals_reduction_launch_and_graph_replay(lib, net):
    """One launch per step (gradient reduction + Adam behind the grid barrier) vs the same kernel followed by the
    reduction launch; then train_resident (chunk CUDA graphs) vs single steps, bit for bit."""
    from deeprank_gnn_b200 import ops
    from deeprank_gnn_b200.engine import Engine
    from test_pinned_path_gpu import _pool
    packed = _pool('cfg3', 8, 32, seed=3, edge_attr=net == 'sGAT')
    kw = dict(device='cuda:0', seed=3, lr=1e-3, graph=True)
    ea, eb, ec = Engine(net, 32, 1, 1, **kw), Engine(net, 32, 1, 1, **kw), Engine(net, 32, 1, 1, **kw)
    eb.fuse_reduce = False
    ra = [ea.upload(pb, slot=i) for i, pb in enumerate(packed)]
    rb = [eb.upload(pb, slot=i) for i, pb in enumerate(packed)]
    rc = [ec.upload(pb, slot=i) for i, pb in enumerate(packed)]
    for i in range(10):
        la, pa = ea.step(ra[i % 8])
        if i < 8:                                   # first use of a slot captures its graph (later steps replay it)
            assert ops.net_step_last()[0] == 1
        lb, pb_ = eb.step(rb[i % 8])
        if i < 8:
            assert ops.net_step_last()[0] == 2
        torch.testing.assert_close(la, lb, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(pa, pb_, rtol=1e-4, atol=1e-5)
    ea.validate(), eb.validate()
    assert float(ea.step_dev[0]) == 10.0 == float(eb.step_dev[0])
    torch.testing.assert_close(ea.params.data, eb.params.data, rtol=1e-3, atol=1e-5)
    lc, pc = ec.train_resident(rc, steps=10)
    ec.validate()
    assert torch.equal(lc, la) and torch.equal(pc, pa) and torch.equal(ec.params.data, ea.params.data)


def test_step3_end_to_end_feeder_for_sgat(lib):
    """Engine.train_batches (C feeder loop: H2D copy, structure-pass graph, step graph, D2H read-back) through
    the general cluster kernel == the same pass issued from Python."""
    from deeprank_gnn_b200.engine import Engine
    from test_pinned_path_gpu import _pool
    packed = _pool('cfg3', 9, 8, seed=11, edge_attr=True)
    ea = Engine('sGAT', 32, 1, 1, device='cuda:0', seed=4, lr=1e-3, graph=True)
    eb = Engine('sGAT', 32, 1, 1, device='cuda:0', seed=4, lr=1e-3, graph=True)
    eb.native_feed = False
    la, pa = ea.train_batches(packed)
    lb, pb_ = eb.train_batches(packed)
    ea.validate(), eb.validate()
    assert ea._feed_keep is not None and eb._feed_keep is None
    assert torch.equal(la, lb) and torch.equal(ea.params.data, eb.params.data)
    for x, y in zip(pa, pb_):
        assert torch.equal(x, y)


@pytest.mark.parametrize('tiles', [0, 2])
def test_step3_three_layer_sgat_matches_oracle(lib, tiles):
    """BASELINE config 3 names "sGAT 3-layer": the throughput variant with a third sGraphAttentionLayer(32, 32) on
    the coarsened graph, fused in the same launch; against the oracle's three-layer restatement."""
    import copy
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.engine import Engine
    from helpers import to_oracle_batch
    from oracle import nets as onets
    from oracle import step as ostep
    graphs = synthetic.make_graphs('cfg3', count=12, seed=6, internal=False)
    torch.manual_seed(3)
    model = onets.sGAT3(32, 1, 1).eval()
    sd0 = copy.deepcopy(model.state_dict())
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    loss, pred = ostep.train_step(model, opt, ostep.make_loss('reg'), to_oracle_batch(graphs))
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    eng = Engine('sGAT', 32, 1, 1, device='cuda:0', layers=3).eval()
    eng.step3_tiles = tiles
    assert sorted(eng.state_dict().keys()) == sorted(sd0.keys())
    eng.load_state_dict(sd0)
    eloss, epred = eng.step(_device_batch(graphs))
    eng.validate()
    assert eng._last_path == 'step3'
    _close_abs(epred.view(-1), pred, 'pred')
    _close_abs(eloss.view(-1), loss.view(-1), 'loss')
    for name, g in eng.named_grads().items():
        _close(g, grads[name], 'grad ' + name)


@pytest.mark.parametrize('tiles', [0, 1, 2])
@pytest.mark.parametrize('net', ['GINet', 'sGAT', 'FoutNet'])
def test_first_aggregation_of_the_structure_pass_is_bit_identical_to_the_step_kernels_phase(lib, net, tiles):
    """conv1's input rows computed by the blob structure pass (``drgnn_structure_io.zin1``, default) against the
    step kernels' own aggregation phase (``pre_agg = False``): same bits in every intermediate the kernels mirror,
    same predictions, loss, gradients and weights over several optimiser steps.  Isolated nodes included (Fout NaN
    rows).  tiles 0 = the kernel the engine picks (the CTA-pair kernel for GINet)."""
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.engine import Engine
    graphs = synthetic.make_graphs(dict(nodes=(5, 200), edges_per_node=5, feat=32), count=19, seed=23)
    g = graphs[3]
    keep = (g.edge_index[0] != 0) & (g.edge_index[1] != 0)        # node 0 of one graph loses every edge
    g.edge_index = g.edge_index[:, keep].contiguous()
    g.edge_attr = g.edge_attr[keep].contiguous()
    d = _device_batch(graphs)
    ea = Engine(net, 32, 1, 1, hidden=HIDDEN, device='cuda:0', seed=7, dropout=0.0, lr=1e-3)
    eb = Engine(net, 32, 1, 1, hidden=HIDDEN, device='cuda:0', seed=7, dropout=0.0, lr=1e-3)
    ea.pre_agg, eb.pre_agg = '1', '0'
    for e in (ea, eb):
        e.keep_intermediates = True
        e.step3_tiles = tiles             # != 0: the general cluster kernel with that tile count, GINet included
    for step in range(3):
        la, pa = ea.step(d)
        lb, pb = eb.step(d)
        ea.validate(), eb.validate()
        assert ea._last_path == eb._last_path
        st = ea._last_struct
        if st.blob_only:      # (the CTA-pair kernel mirrors its intermediates from the full structure pass: nothing precomputed)
            assert st.zin1 is not None and eb._last_struct.zin1 is None
        for name in ('Zin1', 'Z1', 'Zin2', 'Z2'):
            a, b = getattr(ea.ws, name), getattr(eb.ws, name)
            assert torch.equal(torch.nan_to_num(a, nan=-7.0), torch.nan_to_num(b, nan=-7.0)), name
        assert torch.equal(torch.nan_to_num(pa, nan=-7.0), torch.nan_to_num(pb, nan=-7.0))
        assert torch.equal(torch.nan_to_num(la, nan=-7.0), torch.nan_to_num(lb, nan=-7.0))
        for name, gr in ea.named_grads().items():
            assert torch.equal(torch.nan_to_num(gr, nan=-7.0), torch.nan_to_num(eb.named_grads()[name], nan=-7.0)), name
    assert torch.equal(torch.nan_to_num(ea.params.data, nan=-7.0), torch.nan_to_num(eb.params.data, nan=-7.0))
