"""End-to-end parity of the fused engine (forward, loss, backward, Adam) against the CPU
oracle restatement of the reference networks, on the real fixture and on synthetic batches
of the BASELINE shapes.  Tolerance: 1e-4 absolute on outputs and on every parameter
gradient (BASELINE.json north_star), checked relative to the tensor's scale for the large
gradients."""
import copy

import pytest
import torch
import torch.nn.functional as F

from helpers import to_oracle_batch
from oracle import nets as onets
from oracle import step as ostep

pytestmark = pytest.mark.gpu

NETS = {'GINet': onets.GINet, 'sGAT': onets.sGAT, 'FoutNet': onets.FoutNet}


def _fixture_graphs(features=('type', 'polarity', 'bsa'), target='irmsd'):
    from deeprank_gnn_b200.DataSet import HDF5DataSet
    from conftest import FIXTURE
    ds = HDF5DataSet(database=FIXTURE, node_feature=list(features), target=target)
    return [ds.get(i) for i in range(ds.len())]


def _close(a, b, name, tol=1e-4):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, '%s: shape %s vs %s' % (name, tuple(a.shape), tuple(b.shape))
    scale = max(1.0, float(b.abs().max()))
    err = float((a - b).abs().max())
    assert err <= tol * scale, '%s: max|diff| = %.3e (scale %.3e)' % (name, err, scale)


def _oracle_run(net, graphs, hidden, out, task='reg', classes=(0, 1), weights=None, keep_mask=None, train=True, seed=0,
                lr=0.01, literal=True):
    torch.manual_seed(seed)
    onets.LITERAL = literal
    F_in = graphs[0].x.size(1)
    model = NETS[net](F_in, out, 1, hidden=hidden)
    sd0 = copy.deepcopy(model.state_dict())
    model.train(train)
    batch = to_oracle_batch(graphs)
    opt = torch.optim.Adam(model.parameters(), lr=lr)
    loss_fn = ostep.make_loss(task, weights)
    real_dropout = F.dropout
    if keep_mask is not None:
        onets.F.dropout = lambda x, p, training=True: x * keep_mask / (1 - p) if training else x
    try:
        loss, pred = ostep.train_step(model, opt, loss_fn, batch, task, classes)
    finally:
        onets.F.dropout = real_dropout
        onets.LITERAL = True
    grads = {n: p.grad.clone() for n, p in model.named_parameters()}
    return sd0, loss, pred, grads, copy.deepcopy(model.state_dict())


def _engine(net, graphs, hidden, out, sd0, task='reg', weights=None, lr=0.01, **kw):
    from deeprank_gnn_b200.engine import Engine
    eng = Engine(net, graphs[0].x.size(1), out, 1, hidden=hidden, device='cuda:0', task=task, class_weights=weights,
                 lr=lr, **kw)
    eng.load_state_dict(sd0)
    return eng


def _device_batch(graphs, classes=None):
    from deeprank_gnn_b200.data import Batch
    from deeprank_gnn_b200.engine import DeviceBatch
    return DeviceBatch.from_batch(Batch.from_data_list(graphs), 'cuda:0', classes=classes)


@pytest.mark.parametrize('net', ['GINet', 'sGAT', 'FoutNet'])
@pytest.mark.parametrize('data', ['fixture', 'cfg2'])
def test_train_step_matches_oracle(lib, net, data):
    from deeprank_gnn_b200 import synthetic
    if data == 'fixture':
        graphs = _fixture_graphs()[:8]                  # BASELINE config 1: batch of 8, F = 3
    else:
        graphs = synthetic.make_graphs('cfg2', count=6 if net == 'FoutNet' else 16, seed=1)
    hidden = (16, 32)
    keep = None
    if net == 'GINet':
        keep = (torch.rand(len(graphs), 4 * hidden[1], generator=torch.Generator().manual_seed(5)) > 0.4).float()
    sd0, loss, pred, grads, sd1 = _oracle_run(net, graphs, hidden, 1, keep_mask=keep)
    eng = _engine(net, graphs, hidden, 1, sd0)
    d = _device_batch(graphs)
    eloss, epred = eng.step(d, keep_mask=keep)
    eng.validate()
    _close(epred.view(-1), pred, 'pred')
    _close(eloss[0], loss, 'loss')
    for name, g in eng.named_grads().items():
        _close(g, grads[name], 'grad ' + name)
    # Adam at step 1 moves every parameter by ~lr * sign(grad): compare where the gradient is not
    # vanishing (a sign flip of a ~1e-9 gradient is not a parity failure)
    for name, p in eng.state_dict().items():
        solid = grads[name].abs() > 1e-5
        if solid.any():
            _close(p.cpu()[solid], sd1[name][solid], 'param ' + name)


def test_ginet_dead_attention_parameters_keep_zero_grad(lib):
    from deeprank_gnn_b200 import synthetic
    graphs = synthetic.make_graphs('cfg2', count=4, seed=2)
    sd0, loss, pred, grads, sd1 = _oracle_run('GINet', graphs, (16, 32), 1, train=False)
    eng = _engine('GINet', graphs, (16, 32), 1, sd0).eval()
    eng.step(_device_batch(graphs))
    for name, g in eng.named_grads().items():
        if 'attention' in name or 'edge_attr' in name:
            assert float(g.abs().max()) == 0.0 and float(grads[name].abs().max()) == 0.0
            assert torch.equal(eng.state_dict()[name].cpu(), sd0[name])


def test_classification_with_class_weights(lib):
    graphs = _fixture_graphs(target='binclass')
    for i, g in enumerate(graphs):                      # the fixture is all class 0: make it two-class
        g.y = torch.tensor([float(i % 2)])
    w = torch.tensor([0.3, 1.7])
    sd0, loss, pred, grads, sd1 = _oracle_run('GINet', graphs, (16, 32), 2, task='class', weights=w, train=False)
    eng = _engine('GINet', graphs, (16, 32), 2, sd0, task='class', weights=w).eval()
    d = _device_batch(graphs, classes=[0, 1])
    inv = 1.0 / float(w[d.y_class.cpu()].sum())
    eloss, epred = eng.step(d, inv_norm=inv)
    _close(epred, pred, 'logits')
    _close(eloss[0], loss, 'loss')
    for name, g in eng.named_grads().items():
        _close(g, grads[name], 'grad ' + name)


@pytest.mark.parametrize('net,cfg,count', [('GINet', 'cfg4', 4), ('FoutNet', 'cfg5', 3), ('sGAT', 'cfg3', 8)])
def test_wider_configs_match_vectorised_oracle(lib, net, cfg, count):
    """cfg4 (hidden 32/64) and cfg5 (mixed sizes, FoutNet) against the loop-free oracle forms
    (proven equal to the literal loops in tests/test_oracle.py)."""
    from deeprank_gnn_b200 import synthetic
    c = synthetic.CONFIGS[cfg]
    graphs = synthetic.make_graphs(cfg, count=count, seed=4)
    sd0, loss, pred, grads, sd1 = _oracle_run(net, graphs, c['hidden'], 1, train=False, literal=False)
    eng = _engine(net, graphs, c['hidden'], 1, sd0).eval()
    eloss, epred = eng.step(_device_batch(graphs))
    eng.validate()
    _close(epred.view(-1), pred, 'pred')
    _close(eloss[0], loss, 'loss')
    for name, g in eng.named_grads().items():
        _close(g, grads[name], 'grad ' + name)


def test_fout_isolated_node_follows_reference_nan_rule(lib):
    """A node without neighbour gives a NaN FoutLayer row (foutnet.py:73); scatter_max never
    selects a NaN, so the network output and all gradients stay finite and equal the oracle's."""
    from deeprank_gnn_b200 import synthetic
    graphs = synthetic.make_graphs('cfg2', count=3, seed=6)
    g0 = graphs[0]
    victim = 17
    keep = (g0.edge_index[0] != victim) & (g0.edge_index[1] != victim)
    g0.edge_index = g0.edge_index[:, keep].contiguous()
    g0.edge_attr = g0.edge_attr[keep].contiguous()
    sd0, loss, pred, grads, sd1 = _oracle_run('FoutNet', graphs, (16, 32), 1, train=False)
    assert torch.isfinite(pred).all()
    eng = _engine('FoutNet', graphs, (16, 32), 1, sd0).eval()
    eloss, epred = eng.step(_device_batch(graphs))
    _close(epred.view(-1), pred, 'pred')
    for name, g in eng.named_grads().items():
        assert torch.isfinite(g).all(), name
        _close(g, grads[name], 'grad ' + name)


def test_single_graph_batch(lib):
    graphs = _fixture_graphs()[3:4]
    sd0, loss, pred, grads, sd1 = _oracle_run('sGAT', graphs, (16, 32), 1)
    eng = _engine('sGAT', graphs, (16, 32), 1, sd0)
    eloss, epred = eng.step(_device_batch(graphs))
    _close(epred.view(-1), pred, 'pred')
    for name, g in eng.named_grads().items():
        _close(g, grads[name], 'grad ' + name)


@pytest.mark.parametrize('net', ['GINet', 'sGAT'])
def test_packed_graph_replay_equals_eager(lib, net):
    """PackedBatch (int32, one H2D copy) + CUDA-graph replay == eager int64 path, several steps."""
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.data import Batch, PackedBatch
    from deeprank_gnn_b200.engine import Engine
    batches = [synthetic.make_graphs('cfg2', count=8, seed=s) for s in (1, 2, 3)]
    e1 = Engine(net, 32, 1, 1, device='cuda:0', seed=3, dropout=0.0)
    e2 = Engine(net, 32, 1, 1, device='cuda:0', seed=3, dropout=0.0, graph=True)
    e3 = Engine(net, 32, 1, 1, device='cuda:0', seed=3, dropout=0.0, tiled='force')
    for rep in range(2):
        for graphs in batches:
            l1, p1 = e1.step(_device_batch(graphs))
            pb = PackedBatch.from_batch(Batch.from_data_list(graphs))
            l2, p2 = e2.step(e2.upload(pb))
            l3, p3 = e3.step(e3.upload(pb))
            # same kernels, but the packed path sizes level-1 launches by the fixed bound N instead of
            # len(cluster1): the weight-gradient reduction is chunked differently (fp32 summation order)
            torch.testing.assert_close(p2, p1, rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(l2, l1, rtol=1e-4, atol=1e-5)
            torch.testing.assert_close(p3, p1, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(e2.params.data, e1.params.data, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(e3.params.data, e1.params.data, rtol=1e-3, atol=1e-4)


def test_mathmode_tf32x3_within_tolerance(lib):
    from deeprank_gnn_b200 import ops, synthetic
    graphs = synthetic.make_graphs('cfg2', count=8, seed=8)
    sd0, loss, pred, grads, sd1 = _oracle_run('GINet', graphs, (16, 32), 1, train=False)
    eng = _engine('GINet', graphs, (16, 32), 1, sd0).eval()
    old = ops.default_math
    ops.default_math = ops.MATH_TF32X3
    try:
        eloss, epred = eng.step(_device_batch(graphs))
        _close(epred.view(-1), pred, 'pred')
        for name, g in eng.named_grads().items():
            _close(g, grads[name], 'grad ' + name)
    finally:
        ops.default_math = old


@pytest.mark.parametrize('fused_head,variant', [(True, 1), (True, 2), (False, 1)])
def test_fused_per_graph_kernels_equal_op_by_op_path(lib, fused_head, variant):
    """GINet: per-graph fused kernels (csrc/fused.cu) vs the op-by-op launches.  fused_head=True is
    the whole-step kernel (forward + head + loss + backward in one launch; variant 1 = one CTA per
    graph, variant 2 = a cluster of two CTAs per graph, one branch each, everything in shared
    memory), False the fused forward / backward kernels around the op-level head.  Intermediates
    are bit-identical (same summation orders); weight gradients agree to fp32 summation order."""
    from deeprank_gnn_b200 import ops, synthetic
    from deeprank_gnn_b200.engine import Engine
    nodes = (30, 260) if variant == 1 else (5, 200)
    graphs = synthetic.make_graphs(dict(nodes=nodes, edges_per_node=5, feat=32), count=21, seed=13)
    d = _device_batch(graphs)
    e_o = Engine('GINet', 32, 1, 1, device='cuda:0', seed=5, dropout=0.0, fused_graph=False, fused_head=False)
    e_f = Engine('GINet', 32, 1, 1, device='cuda:0', seed=5, dropout=0.0, fused_head=fused_head)
    e_f.step_variant, e_f.keep_intermediates = variant, True
    e_f.fused_tc = False                  # fp32 FMA tiles: the intermediates are then bit-identical to the op-level kernels
    assert e_f._use_fused_graph(d) and not e_o._use_fused_graph(d)
    for step in range(3):
        lo, po = e_o.step(d)
        lf, pf = e_f.step(d)
        e_f.validate()
        if fused_head:
            assert ops.ginet_step_last_variant() == variant
        if step == 0:
            N, (K0, _E1, K1) = d.N, e_o.structs[0].sync_counts()
            for name in ('Zin1', 'Z1'):
                assert torch.equal(getattr(e_f.ws, name)[:N], getattr(e_o.ws, name)[:N]), name
            for name in ('arg0', 'Zin2', 'Z2'):
                assert torch.equal(getattr(e_f.ws, name)[:K0], getattr(e_o.ws, name)[:K0]), name
            assert torch.equal(e_f.ws.arg1[:K1], e_o.ws.arg1[:K1])
            torch.testing.assert_close(e_f.ws.R[:d.B], e_o.ws.R[:d.B], rtol=1e-6, atol=1e-7)
        torch.testing.assert_close(pf, po, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(lf, lo, rtol=1e-4, atol=1e-6)
        gf, go = e_f.named_grads(), e_o.named_grads()
        for name in gf:
            torch.testing.assert_close(gf[name], go[name], rtol=1e-3, atol=1e-5, msg=name)


def test_whole_step_kernel_classification_and_dropout(lib):
    """Whole-step kernel with class weights + injected dropout mask vs the op-by-op path."""
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.engine import Engine
    graphs = synthetic.make_graphs('cfg2', count=12, seed=21)
    for i, g in enumerate(graphs):
        g.y = torch.tensor([float(i % 3)])
    w = torch.tensor([0.5, 1.5, 1.0])
    d = _device_batch(graphs, classes=[0, 1, 2])
    inv = 1.0 / float(w[d.y_class.cpu()].sum())
    keep = (torch.rand(12, 128, generator=torch.Generator().manual_seed(1)) > 0.4).float()
    kw = dict(device='cuda:0', seed=9, task='class', class_weights=w)
    e_o = Engine('GINet', 32, 3, 1, fused_graph=False, fused_head=False, **kw)
    e_f = Engine('GINet', 32, 3, 1, **kw)
    lo, po = e_o.step(d, inv_norm=inv, keep_mask=keep)
    lf, pf = e_f.step(d, inv_norm=inv, keep_mask=keep)
    assert e_f._all_done
    torch.testing.assert_close(pf, po, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(lf, lo, rtol=1e-4, atol=1e-6)
    for name, gfv in e_f.named_grads().items():
        torch.testing.assert_close(gfv, e_o.named_grads()[name], rtol=1e-3, atol=1e-5, msg=name)
    # forward only (scoring) through the same kernel
    pe = e_f.eval().forward(d)
    po2 = e_o.eval().forward(d)
    torch.testing.assert_close(pe, po2, rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize('task', ['reg', 'class'])
def test_cluster_step_kernel_equals_single_cta_kernel(lib, task):
    """The two whole-step kernels on the same batch: predictions / loss to fp32 rounding (the graph
    part keeps every summation order - see the bit-identical intermediates in
    test_fused_per_graph_kernels_equal_op_by_op_path - the head sums fc1 differently), gradients to
    fp32 summation order (split-K weight gradients), hashed dropout identical, several optimiser
    steps stay together."""
    from deeprank_gnn_b200 import ops, synthetic
    from deeprank_gnn_b200.engine import Engine
    graphs = synthetic.make_graphs('cfg2', count=16, seed=77)
    kw = dict(device='cuda:0', seed=11, lr=1e-3)
    classes = None
    if task == 'class':
        for i, g in enumerate(graphs):
            g.y = torch.tensor([float(i % 2)])
        kw.update(task='class', class_weights=torch.tensor([0.7, 1.3]))
        classes = [0, 1]
    d = _device_batch(graphs, classes=classes)
    out = 2 if task == 'class' else 1
    inv = None if task == 'reg' else 1.0 / float(kw['class_weights'][d.y_class.cpu()].sum())
    e1 = Engine('GINet', 32, out, 1, **kw)
    e2 = Engine('GINet', 32, out, 1, **kw)
    e1.step_variant, e2.step_variant = 1, 2
    for step in range(4):
        l1, p1 = e1.step(d, inv_norm=inv)
        assert ops.ginet_step_last_variant() == 1
        l2, p2 = e2.step(d, inv_norm=inv)
        assert ops.ginet_step_last_variant() == 2
        e1.validate(), e2.validate()
        if step == 0:
            # same forward arithmetic; the cluster kernel sums fc1 in a different (fixed) order
            torch.testing.assert_close(p1, p2, rtol=1e-5, atol=1e-6)
            torch.testing.assert_close(l1, l2, rtol=1e-5, atol=1e-6)
            g1, g2 = e1.named_grads(), e2.named_grads()
            for name in g1:
                torch.testing.assert_close(g2[name], g1[name], rtol=1e-3, atol=1e-6, msg=name)
        torch.testing.assert_close(p2, p1, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(e2.params.data, e1.params.data, rtol=1e-3, atol=1e-4)
    # scoring (forward only) through the cluster kernel
    s1, s2 = e1.eval().forward(d), e2.eval().forward(d)
    torch.testing.assert_close(s2, s1, rtol=1e-3, atol=1e-4)


def test_cluster_step_kernel_falls_back_when_a_graph_does_not_fit(lib):
    from deeprank_gnn_b200 import ops, synthetic
    from deeprank_gnn_b200.engine import Engine
    from deeprank_gnn_b200._lib import DrgnnError
    graphs = synthetic.make_graphs(dict(nodes=(250, 300), edges_per_node=5, feat=32), count=4, seed=3)
    d = _device_batch(graphs)
    assert ops.ginet_step2_smem_bytes(32, 16, 32, d.max_n, d.max_k0, d.max_k1, d.max_e, 128, 1) < 0
    e = Engine('GINet', 32, 1, 1, device='cuda:0', seed=1, dropout=0.0)
    e.step3 = False                       # without the general cluster kernel: the single-CTA whole-step kernel
    e.step(d)
    assert ops.ginet_step_last_variant() == 1 and e._last_path == 'ops'
    e.step_variant = 2
    with pytest.raises(DrgnnError):
        e.step(d)
    e2 = Engine('GINet', 32, 1, 1, device='cuda:0', seed=1, dropout=0.0)
    e2.load_state_dict(e.state_dict())
    e.step_variant = 1
    l1, p1 = e.step(d)
    l2, p2 = e2.step(d)                   # default: the general cluster kernel (one branch per CTA, as many node tiles as needed)
    assert e2._last_path == 'step3' and ops.net_step_last()[1] >= 1
    torch.testing.assert_close(p2, p1, rtol=1e-4, atol=1e-5)


def test_cluster_step_kernel_in_kernel_reduction_equals_reduction_launch(lib):
    """Cluster kernel with the gradient reduction + Adam behind a grid barrier (one launch per step)
    vs the same kernel followed by the reduction launch: same gradients to fp32 summation order, the
    optimiser state stays together over several steps, and a batch too large to be co-resident takes
    the two-launch path."""
    from deeprank_gnn_b200 import _lib, ops, synthetic
    from deeprank_gnn_b200.engine import Engine
    graphs = synthetic.make_graphs('cfg2', count=24, seed=5)
    d = _device_batch(graphs)
    ea = Engine('GINet', 32, 1, 1, device='cuda:0', seed=3, lr=1e-3)
    eb = Engine('GINet', 32, 1, 1, device='cuda:0', seed=3, lr=1e-3)
    ea.step_variant = eb.step_variant = 2
    eb.fuse_reduce = False
    for step in range(5):
        la, pa = ea.step(d)
        assert _lib.load().drgnn_ginet_step_last_launches() == 1
        lb, pb = eb.step(d)
        assert _lib.load().drgnn_ginet_step_last_launches() == 2
        ea.validate(), eb.validate()
        torch.testing.assert_close(la, lb, rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(pa, pb, rtol=1e-4, atol=1e-5)
        if step == 0:
            for name, g in ea.named_grads().items():
                torch.testing.assert_close(g, eb.named_grads()[name], rtol=1e-4, atol=1e-6, msg=name)
    assert float(ea.step_dev[0]) == 5.0 and float(eb.step_dev[0]) == 5.0
    torch.testing.assert_close(ea.params.data, eb.params.data, rtol=1e-3, atol=1e-5)
    big = _device_batch(synthetic.make_graphs('cfg2', count=96, seed=6))
    ea.step(big)
    assert ops.ginet_step_last_variant() == 2 and _lib.load().drgnn_ginet_step_last_launches() == 2
    ea.validate()


def test_compact_packed_batches_train_like_reference_dtype_batches(lib):
    """PackedBatch(idx16=True, edge_attr=False) - uint16 graph-local edge ids, no edge attributes (GINet's
    attention is the identity) - trains exactly like the int64 reference-dtype batch."""
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.data import Batch, PackedBatch
    from deeprank_gnn_b200.engine import Engine
    graphs = synthetic.make_graphs(dict(nodes=(20, 200), edges_per_node=5, feat=32), count=12, seed=31)
    batch = Batch.from_data_list(graphs)
    ea = Engine('GINet', 32, 1, 1, device='cuda:0', seed=2, dropout=0.0)
    eb = Engine('GINet', 32, 1, 1, device='cuda:0', seed=2, dropout=0.0)
    da = _device_batch(graphs)
    pb = PackedBatch.from_batch(batch, idx16=True, edge_attr=False)
    assert pb.nbytes < PackedBatch.from_batch(batch).nbytes
    db = eb.upload(pb)
    for _ in range(3):
        la, pa = ea.step(da)
        lb, pred_b = eb.step(db)
        ea.validate(), eb.validate()
        assert torch.equal(pa, pred_b) and torch.equal(la, lb)
    assert torch.equal(ea.params.data, eb.params.data)


def test_native_feed_loop_equals_python_pipeline(lib):
    """Engine.train_batches with the loop issued from C (drgnn_feed_run: H2D copy, structure-pass graph,
    step graph, D2H read-back per step) against the same pass issued from Python: identical losses,
    predictions and weights."""
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.data import Batch, PackedBatch
    from deeprank_gnn_b200.engine import Engine
    packed = []
    for i in range(11):
        graphs = synthetic.make_graphs('cfg2', count=6, seed=100 + i)
        packed.append(PackedBatch.from_batch(Batch.from_data_list(graphs), idx16=True, edge_attr=False))
    ea = Engine('GINet', 32, 1, 1, device='cuda:0', seed=4, lr=1e-3, graph=True)
    eb = Engine('GINet', 32, 1, 1, device='cuda:0', seed=4, lr=1e-3, graph=True)
    eb.native_feed = False
    la, pa = ea.train_batches(packed)
    lb, pb_ = eb.train_batches(packed)
    ea.validate(), eb.validate()
    assert ea._feed_keep is not None and eb._feed_keep is None
    assert torch.equal(la, lb)
    for x, y in zip(pa, pb_):
        assert torch.equal(x, y)
    assert torch.equal(ea.params.data, eb.params.data)
    assert float(ea.step_dev[0]) == 11.0


def test_cluster_step_kernel_edge_cases_match_oracle(lib):
    """SURVEY 8a-ter through the cluster kernel + bitmap structure pass, against the CPU oracle: a single
    graph (B = 1: get_preloaded_cluster's loop never runs; the in-kernel reduction sweeps 5 k elements per
    CTA), a graph whose clusters swallow every edge (pooled edge_index empty: conv2 aggregates nothing),
    singleton clusters with gaps in the ids (consecutive_cluster closes them), an isolated node (deg 0 ->
    zero row) and duplicate edges (aggregated twice at level 0, coalesced at level 1)."""
    from deeprank_gnn_b200 import ops, synthetic
    graphs = synthetic.make_graphs(dict(nodes=(40, 120), edges_per_node=5, feat=32), count=4, seed=41)
    g = graphs[1]
    n = g.x.size(0)
    # all edges intra-cluster: two clusters = the two connected halves would still have cross edges, so use ONE cluster
    g.cluster0 = torch.zeros(n, dtype=torch.long)
    g.cluster1 = torch.zeros(1, dtype=torch.long)
    h = graphs[2]
    nh = h.x.size(0)
    h.cluster0 = torch.arange(nh, dtype=torch.long) * 3 + 7          # singletons, ids with gaps
    h.cluster1 = (torch.arange(nh, dtype=torch.long) // 2) * 5       # pairs, ids with gaps
    k = graphs[3]
    victim = 5
    keep = (k.edge_index[0] != victim) & (k.edge_index[1] != victim)
    k.edge_index = torch.cat([k.edge_index[:, keep], k.edge_index[:, keep][:, :7]], dim=1).contiguous()   # + duplicates
    k.edge_attr = torch.cat([k.edge_attr[keep], k.edge_attr[keep][:7]]).contiguous()
    for subset in (graphs, graphs[1:2], graphs[:1]):
        sd0, loss, pred, grads, sd1 = _oracle_run('GINet', subset, (16, 32), 1, train=False)
        eng = _engine('GINet', subset, (16, 32), 1, sd0).eval()
        eng.step_variant = 2
        d = _device_batch(subset)
        eloss, epred = eng.step(d)
        eng.validate()
        assert ops.ginet_step_last_variant() == 2 and eng.structs[d.sslot].blob_only
        _close(epred.view(-1), pred, 'pred')
        _close(eloss.view(-1), loss.view(-1), 'loss')
        for name, gr in eng.named_grads().items():
            _close(gr, grads[name], 'grad ' + name)


def test_native_feed_scoring_equals_python_pipeline(lib):
    """Forward-only pass (NeuralNet.eval / test) through the C feeder loop: same losses and predictions as
    the Python loop, weights untouched."""
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.data import Batch, PackedBatch
    from deeprank_gnn_b200.engine import Engine
    packed = []
    for i in range(9):
        graphs = synthetic.make_graphs('cfg2', count=5, seed=300 + i)
        packed.append(PackedBatch.from_batch(Batch.from_data_list(graphs), idx16=True, edge_attr=False))
    ea = Engine('GINet', 32, 1, 1, device='cuda:0', seed=8, graph=True)
    eb = Engine('GINet', 32, 1, 1, device='cuda:0', seed=8, graph=True)
    eb.native_feed = False
    w0 = ea.params.data.clone()
    la, pa = ea.train_batches(packed, train=False)
    lb, pb_ = eb.train_batches(packed, train=False)
    assert ea._feed_keep is not None and eb._feed_keep is None
    torch.testing.assert_close(la, lb, rtol=1e-6, atol=1e-7)
    for x, y in zip(pa, pb_):
        assert torch.equal(x, y)
    assert torch.equal(ea.params.data, w0) and float(ea.step_dev[0]) == 0.0


def test_cluster_step_kernel_tensor_core_tiles_match_fma_tiles(lib):
    """The dense products of the cluster kernel on mma.sync 3xTF32 tiles (opt-in) against the fp32 FMA register tiles:
    same predictions, loss and gradients to 1e-5, several optimiser steps stay together."""
    from deeprank_gnn_b200 import ops, synthetic
    from deeprank_gnn_b200.engine import Engine
    graphs = synthetic.make_graphs(dict(nodes=(20, 200), edges_per_node=5, feat=32), count=24, seed=9)
    d = _device_batch(graphs)
    ea = Engine('GINet', 32, 1, 1, device='cuda:0', seed=3, lr=1e-3, dropout=0.0)
    eb = Engine('GINet', 32, 1, 1, device='cuda:0', seed=3, lr=1e-3, dropout=0.0)
    ea.step_variant = eb.step_variant = 2
    ea.fused_tc, eb.fused_tc = True, False
    for step in range(4):
        la, pa = ea.step(d)
        lb, pb = eb.step(d)
        assert ops.ginet_step_last_variant() == 2
        ea.validate(), eb.validate()
        torch.testing.assert_close(pa, pb, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(la, lb, rtol=1e-5, atol=1e-6)
        if step == 0:
            for name, g in ea.named_grads().items():
                torch.testing.assert_close(g, eb.named_grads()[name], rtol=1e-4, atol=1e-5, msg=name)
    torch.testing.assert_close(ea.params.data, eb.params.data, rtol=1e-3, atol=1e-5)


@pytest.mark.parametrize('dropout', [0.0, 0.4])
@pytest.mark.parametrize('fuse_reduce', [True, False])
def test_cluster_step_kernel_head_v2_is_bit_identical(lib, dropout, fuse_reduce):
    """Head v2 of the CTA-pair kernel (flag bit 6: fc2 / loss / dLoss/dpred in every warp, fc1.weight gradient formed
    inside the in-kernel reduction from the fc1.bias gradient rows and the read-out rows instead of stored rows)
    against the first version: predictions, loss, every gradient, the weights and both Adam moments after several
    steps are the SAME BITS (rounded products, same summation order), with and without the in-kernel reduction."""
    from deeprank_gnn_b200 import ops, synthetic
    from deeprank_gnn_b200.engine import Engine
    graphs = synthetic.make_graphs(dict(nodes=(20, 200), edges_per_node=5, feat=32), count=40, seed=21)
    d = _device_batch(graphs)
    ea = Engine('GINet', 32, 1, 1, device='cuda:0', seed=3, lr=1e-3, dropout=dropout)
    eb = Engine('GINet', 32, 1, 1, device='cuda:0', seed=3, lr=1e-3, dropout=dropout)
    ea.step_variant = eb.step_variant = 2
    ea.fuse_reduce = eb.fuse_reduce = fuse_reduce
    ea.head_v2, eb.head_v2 = True, False
    la, pa = ea.loss_and_grads(d)
    lb, pb = eb.loss_and_grads(d)
    assert ops.ginet_step_last_variant() == 2
    assert torch.equal(pa, pb) and torch.equal(la, lb)
    for name, g in ea.named_grads().items():
        assert torch.equal(g, eb.named_grads()[name]), name
    for _ in range(4):
        la, pa = ea.step(d)
        lb, pb = eb.step(d)
        ea.validate(), eb.validate()
        assert torch.equal(pa, pb) and torch.equal(la, lb)
    assert torch.equal(ea.params.data, eb.params.data)
    assert torch.equal(ea.exp_avg, eb.exp_avg) and torch.equal(ea.exp_avg_sq, eb.exp_avg_sq)
