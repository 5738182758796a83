"""GPU parity of the drop-in nn.Module API (GINet / sGAT / FoutNet, the stand-alone layers,
community_pooling, get_preloaded_cluster) and of the NeuralNet driver against the CPU oracle."""
import copy
import os

import numpy as np
import pytest
import torch

from conftest import FIXTURE
from helpers import to_oracle_batch, to_oracle_data
from oracle import nets as onets
from oracle import pooling as opool
from oracle import pyg_min

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _close(a, b, name, tol=1e-4):
    a, b = a.detach().cpu().double(), b.detach().cpu().double()
    assert a.shape == b.shape, '%s: %s vs %s' % (name, tuple(a.shape), tuple(b.shape))
    scale = max(1.0, float(b.abs().max()))
    err = float((a - b).abs().max())
    assert err <= tol * scale, '%s: max|diff| %.3e (scale %.3e)' % (name, err, scale)


def _graphs(kind):
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.DataSet import HDF5DataSet
    if kind == 'fixture':
        ds = HDF5DataSet(database=FIXTURE, node_feature=['type', 'polarity', 'bsa'], target='irmsd')
        return [ds.get(i) for i in range(8)]
    return synthetic.make_graphs('cfg2', count=6, seed=12)


@pytest.mark.parametrize('net', ['GINet', 'sGAT', 'FoutNet'])
@pytest.mark.parametrize('data', ['fixture', 'cfg2'])
def test_module_forward_backward_matches_oracle(lib, net, data):
    from deeprank_gnn_b200 import foutnet, ginet, sGAT
    from deeprank_gnn_b200.data import Batch
    graphs = _graphs(data)
    F_in = graphs[0].x.size(1)
    torch.manual_seed(0)
    ref = {'GINet': onets.GINet, 'sGAT': onets.sGAT, 'FoutNet': onets.FoutNet}[net](F_in, 1, 1).eval()
    mod = {'GINet': ginet.GINet, 'sGAT': sGAT.sGAT, 'FoutNet': foutnet.FoutNet}[net](F_in, 1, 1).to(DEV).eval()
    assert list(mod.state_dict().keys()) == list(ref.state_dict().keys())
    mod.load_state_dict(ref.state_dict())
    ob = to_oracle_batch(graphs)
    y = ob.y.clone()
    pred_ref = ref(ob).reshape(-1)
    torch.nn.MSELoss()(pred_ref, y).backward()
    batch = Batch.from_data_list(graphs).to(DEV)
    pred = mod(batch).reshape(-1)
    torch.nn.MSELoss()(pred, y.to(DEV)).backward()
    _close(pred, pred_ref, 'pred')
    refg = dict(ref.named_parameters())
    for name, p in mod.named_parameters():
        assert p.grad is not None, name
        _close(p.grad, refg[name].grad, 'grad ' + name)


def test_ginet_train_mode_dropout_and_dead_params(lib):
    from deeprank_gnn_b200 import ginet
    from deeprank_gnn_b200.data import Batch
    graphs = _graphs('cfg2')
    mod = ginet.GINet(32, 1, 1).to(DEV).train()
    out = mod(Batch.from_data_list(graphs).to(DEV))
    out.sum().backward()
    assert out.shape == (6, 1) and torch.isfinite(out).all()
    for name, p in mod.named_parameters():
        assert p.grad is not None
        if 'attention' in name or 'edge_attr' in name:
            assert float(p.grad.abs().max()) == 0.0


@pytest.mark.parametrize('layer', ['ginet', 'sgat', 'fout'])
def test_standalone_layers_match_oracle(lib, layer):
    from deeprank_gnn_b200 import foutnet, ginet, sGAT
    g = torch.Generator().manual_seed(3)
    n, E, Fi, Fo = 60, 300, 12, 8
    x = torch.randn(n, Fi, generator=g)
    ei = torch.randint(0, n, (2, E), generator=g)
    ea = torch.rand(E, 1, generator=g) + 0.5
    torch.manual_seed(1)
    if layer == 'ginet':
        ref, mod = onets.GINetConvLayer(Fi, Fo), ginet.GINetConvLayer(Fi, Fo)
        call = lambda m, xx, e, a: m(xx, e, a)
    elif layer == 'sgat':
        ref, mod = onets.sGraphAttentionLayer(Fi, Fo), sGAT.sGraphAttentionLayer(Fi, Fo)
        call = lambda m, xx, e, a: m(xx, e, a)
    else:
        ei[0, :n] = torch.arange(n)             # every node has a neighbour (else NaN rows, tested elsewhere)
        ref, mod = onets.FoutLayer(Fi, Fo), foutnet.FoutLayer(Fi, Fo)
        call = lambda m, xx, e, a: m(xx, e)
    mod.load_state_dict(ref.state_dict())
    mod = mod.to(DEV)
    xr = x.clone().requires_grad_(True)
    out_ref = call(ref, xr, ei, ea)
    out_ref.square().sum().backward()
    xd = x.clone().to(DEV).requires_grad_(True)
    out = call(mod, xd, ei.to(DEV), ea.to(DEV))
    out.square().sum().backward()
    _close(out, out_ref, 'out')
    _close(xd.grad, xr.grad, 'dx')
    refg = dict(ref.named_parameters())
    for name, p in mod.named_parameters():
        _close(p.grad, refg[name].grad, 'grad ' + name)


def test_community_pooling_api_matches_oracle(lib):
    from deeprank_gnn_b200.community_pooling import community_pooling, get_preloaded_cluster
    from deeprank_gnn_b200.data import Batch
    graphs = _graphs('fixture')
    ob = to_oracle_batch(graphs)
    cl_ref = opool.get_preloaded_cluster(ob.cluster0.clone(), ob.batch)
    xr = ob.x.clone().requires_grad_(True)
    ob.x = xr
    pooled_ref = opool.community_pooling(cl_ref, ob)
    batch = Batch.from_data_list(graphs).to(DEV)
    cl = get_preloaded_cluster(batch.cluster0.clone(), batch.batch)
    assert torch.equal(cl.cpu(), cl_ref)
    xd = batch.x.clone().requires_grad_(True)
    batch.x = xd
    pooled = community_pooling(cl, batch)
    assert torch.equal(pooled.edge_index.cpu(), pooled_ref.edge_index)                  # bit-exact
    assert torch.equal(pooled.batch.cpu(), pooled_ref.batch)
    assert torch.equal(pooled.internal_edge_index.cpu(), pooled_ref.internal_edge_index)
    assert torch.equal(pooled.x.detach().cpu(), pooled_ref.x.detach())
    _close(pooled.edge_attr, pooled_ref.edge_attr, 'edge_attr', 1e-5)
    _close(pooled.internal_edge_attr, pooled_ref.internal_edge_attr, 'internal_edge_attr', 1e-5)
    _close(pooled.pos, pooled_ref.pos, 'pos', 1e-5)
    assert torch.equal(pooled.cluster1.cpu(), pooled_ref.cluster1)
    w = torch.randn(pooled_ref.x.shape)
    (pooled_ref.x * w).sum().backward()
    (pooled.x * w.to(DEV)).sum().backward()
    assert torch.equal(xd.grad.cpu(), xr.grad)


def test_community_pooling_toy_of_reference_test(lib):
    """tests/test_community_pooling.py:52-58: Batch without cluster0 / edge_attr."""
    from deeprank_gnn_b200.community_pooling import community_pooling
    from deeprank_gnn_b200.data import Batch, Data
    ei = torch.tensor([[0, 1, 1, 2, 3, 4, 4, 5], [1, 0, 2, 1, 4, 3, 5, 4]], dtype=torch.long)
    x = torch.tensor([[0.], [1.], [2.], [3.], [4.], [5.]])
    mk = lambda: Data(x=x.clone(), edge_index=ei.clone(), pos=torch.randn(6, 3))
    batch = Batch.from_data_list([mk(), mk()]).to(DEV)
    out = community_pooling(torch.tensor([0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3], device=DEV), batch)
    assert out.x.view(-1).tolist() == [2., 5., 2., 5.] and out.edge_index.numel() == 0
    assert out.batch.tolist() == [0, 0, 1, 1]
    out = community_pooling(torch.tensor([0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5], device=DEV),
                            Batch.from_data_list([mk(), mk()]).to(DEV))
    assert out.edge_index.tolist() == [[0, 1, 1, 2, 3, 4, 4, 5], [1, 0, 2, 1, 4, 3, 5, 4]]
    single = community_pooling(torch.tensor([0, 0, 1, 1, 2, 2], device=DEV), mk().to(DEV))
    assert single.x.view(-1).tolist() == [1., 3., 5.] and not hasattr(single, 'batch')


@pytest.mark.parametrize('net,task', [('GINet', 'reg'), ('GINet', 'class'), ('FoutNet', 'reg'), ('sGAT', 'reg')])
@pytest.mark.parametrize('fused', [True, False])
def test_neuralnet_workloads_of_reference_test_nn(lib, tmp_path, net, task, fused):
    """The four workloads of tests/test_nn.py:9-51 (7 node features, batch 64, 80/20 split, 5 epochs,
    validate, save, reload pretrained) through both execution paths."""
    from deeprank_gnn_b200 import foutnet, ginet, sGAT
    from deeprank_gnn_b200.NeuralNet import NeuralNet
    Net = {'GINet': ginet.GINet, 'sGAT': sGAT.sGAT, 'FoutNet': foutnet.FoutNet}[net]
    feats = ['type', 'polarity', 'bsa', 'depth', 'hse', 'ic', 'pssm']
    kw = dict(target='irmsd') if task == 'reg' else dict(target='binclass', task='class')
    nn_ = NeuralNet(FIXTURE, Net, node_feature=feats, edge_feature=['dist'], batch_size=64, percent=[0.8, 0.2],
                    outdir=str(tmp_path), fused=fused, verbose=False, **kw)
    assert (nn_.engine is not None) == fused
    nn_.train(nepoch=5, validate=True, save_model='last', hdf5='train_data.hdf5')
    assert len(nn_.train_loss) == 5 and all(l == l for l in nn_.train_loss)
    assert len(nn_.train_out) == 8 and len(nn_.valid_out) == 2
    ck = os.path.join(str(tmp_path), 'ck.pth.tar')
    nn_.save_model(ck)
    state = torch.load(ck, weights_only=False)
    assert sorted(state.keys()) == sorted(['model', 'optimizer', 'node', 'edge', 'target', 'task', 'classes',
                                           'class_weight', 'batch_size', 'percent', 'lr', 'index', 'shuffle',
                                           'threshold', 'cluster_nodes', 'transform_sigmoid'])
    nn2 = NeuralNet(FIXTURE, Net, pretrained_model=ck, outdir=str(tmp_path), fused=fused, verbose=False)
    nn2.test(hdf5='test_data.hdf5')
    assert len(nn2.test_out) == 10
    nn_.model.eval()
    for k, v in nn2.model.state_dict().items():
        assert torch.equal(v.cpu(), state['model'][k].cpu())


def test_neuralnet_fused_epoch_equals_autograd_epoch(lib, tmp_path):
    """One epoch (no shuffle, sGAT: no dropout) through the fused engine == the nn.Module/autograd loop."""
    from deeprank_gnn_b200 import sGAT
    from deeprank_gnn_b200.NeuralNet import NeuralNet
    res = []
    for fused in (True, False):
        torch.manual_seed(0)
        np.random.seed(0)                       # DivideDataSet shuffles with the numpy RNG (DataSet.py:27-28)
        nn_ = NeuralNet(FIXTURE, sGAT.sGAT, target='irmsd', batch_size=4, percent=[1.0, 0.0], shuffle=False,
                        outdir=str(tmp_path), fused=fused, verbose=False, lr=0.001)
        nn_.train(nepoch=2, save_model=None, hdf5='cmp.hdf5')
        res.append((nn_.train_loss, {k: v.cpu() for k, v in nn_.model.state_dict().items()}))
    for a, b in zip(res[0][0], res[1][0]):
        assert abs(a - b) <= 1e-3 * max(1.0, abs(b))
    for k in res[0][1]:
        _close(res[0][1][k], res[1][1][k], k, 2e-3)


@pytest.mark.parametrize('net', ['GINet', 'sGAT'])
def test_neuralnet_epochs_from_the_packed_cache_equal_epochs_from_hdf5(lib, tmp_path, net):
    """SURVEY 8f rank 1 wired into the plugin API: ``NeuralNet(..., cache=dir)`` writes the packed feeder records
    once (data.PackedCache: one memory-mapped file, page-locked with cudaHostRegister) and feeds every later epoch
    from the mapping; training from the cache must equal training from ``HDF5DataSet`` batch for batch
    (DataSet.py:231-366 replaced), and a second NeuralNet must reuse the files without touching HDF5 collation."""
    import os
    from deeprank_gnn_b200 import ginet, sGAT
    from deeprank_gnn_b200.NeuralNet import NeuralNet
    Net = {'GINet': ginet.GINet, 'sGAT': sGAT.sGAT}[net]
    kw = dict(node_feature=['type', 'polarity', 'bsa'], edge_feature=['dist'], target='irmsd', lr=0.01, batch_size=3,
              percent=[1.0, 0.0], shuffle=False, outdir=str(tmp_path), verbose=False)
    torch.manual_seed(3)
    np.random.seed(11)                      # DivideDataSet shuffles the graph order with numpy's global RNG
    a = NeuralNet(FIXTURE, Net, **kw)
    sd0 = {k: v.clone() for k, v in a.model.state_dict().items()}
    np.random.seed(11)
    b = NeuralNet(FIXTURE, Net, cache=str(tmp_path / 'cache'), **kw)
    assert a.train_loader.dataset.index_complexes == b.train_loader.dataset.index_complexes
    for m in (a, b):
        m.model.load_state_dict(sd0)
        m.engine.load_state_dict(sd0)
        m.engine.spec.dropout = 0.0
    a.train(nepoch=3, validate=False, save_model='none')
    b.train(nepoch=3, validate=False, save_model='none')
    assert a.train_loss == b.train_loss
    for k, v in a.model.state_dict().items():
        assert torch.equal(v, b.model.state_dict()[k]), k
    files = sorted(os.listdir(str(tmp_path / 'cache')))
    assert len(files) == 2 and files[0].endswith('.json') and files[1].endswith('.pack')
    cache = b._records[id(b.train_loader)][0]
    assert cache.registered or cache.pin           # page-locked mapping (or the pinned-copy fallback)
    stamp = os.path.getmtime(os.path.join(str(tmp_path / 'cache'), files[1]))
    np.random.seed(11)
    c = NeuralNet(FIXTURE, Net, cache=str(tmp_path / 'cache'), **kw)
    c.model.load_state_dict(sd0)
    c.engine.load_state_dict(sd0)
    c.engine.spec.dropout = 0.0
    c.train(nepoch=3, validate=False, save_model='none')
    assert c.train_loss == a.train_loss
    assert os.path.getmtime(os.path.join(str(tmp_path / 'cache'), files[1])) == stamp      # reused, not rebuilt


def test_two_graph_ginet_of_the_documentation_matches_oracle(lib):
    """docs/tutorial.advanced.rst:126-137: interface edges for conv1 / conv2, INTERNAL edges for the ``_ext``
    branch; forward and every parameter gradient against the oracle restatement, on the fixture."""
    from deeprank_gnn_b200 import ginet
    from deeprank_gnn_b200.data import Batch
    graphs = _graphs('fixture')[:6]
    torch.manual_seed(2)
    ref = onets.GINetInternal(graphs[0].x.size(1), 1, 1).eval()
    sd = copy.deepcopy(ref.state_dict())
    pred = ref(to_oracle_batch(graphs))
    pred.pow(2).mean().backward()
    mod = ginet.GINetInternal(graphs[0].x.size(1), 1, 1).to(DEV).eval()
    mod.load_state_dict(sd)
    out = mod(Batch.from_data_list(graphs).to(DEV))
    out.pow(2).mean().backward()
    _close(out, pred, 'pred')
    ref_g = dict(ref.named_parameters())
    for n, p in mod.named_parameters():
        _close(p.grad, ref_g[n].grad, 'grad ' + n)
    # the branches really see different graphs: the shipped single-graph GINet gives another answer
    single = ginet.GINet(graphs[0].x.size(1), 1, 1).to(DEV).eval()
    single.load_state_dict(sd)
    assert float((single(Batch.from_data_list(graphs).to(DEV)) - out).detach().abs().max()) > 1e-4


def test_ginet_conv_layer_with_bias_adds_it_once_per_incoming_edge(lib):
    """GINetConvLayer(bias=True): the reference applies fc (with bias) to x[col] per edge before scatter_sum
    (ginet.py:57,71), so node i receives deg_i * b."""
    from deeprank_gnn_b200 import ginet
    g = _graphs('cfg2')[0]
    torch.manual_seed(4)
    ref = onets.GINetConvLayer(32, 16, 1, bias=True)
    mod = ginet.GINetConvLayer(32, 16, 1, bias=True).to(DEV)
    mod.load_state_dict(ref.state_dict())
    out_ref = ref(g.x, g.edge_index, g.edge_attr)
    out = mod(g.x.to(DEV), g.edge_index.to(DEV), g.edge_attr.to(DEV))
    _close(out, out_ref, 'conv with bias')
