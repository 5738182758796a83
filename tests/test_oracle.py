"""CPU tests that PIN the oracle: the hand-computable toy graph of the reference's own test
(tests/test_community_pooling.py:10-18, 52-58), structural facts of the shipped fixture
(SURVEY 8a), algebraic identities of the reference code, the published semantics of the
un-vendored torch_scatter / torch_geometric primitives, and the shipped checkpoints' layout."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import FIXTURE, ROOT
from helpers import oracle_structure, to_oracle_batch
from oracle import nets as onets
from oracle import pooling, pyg_min


def _toy():
    ei = torch.tensor([[0, 1, 1, 2, 3, 4, 4, 5], [1, 0, 2, 1, 4, 3, 5, 4]], dtype=torch.long)
    x = torch.tensor([[0.], [1.], [2.], [3.], [4.], [5.]])
    return pyg_min.Data(x=x, edge_index=ei, pos=torch.randn(6, 3))


def test_toy_graph_pooling_golden():
    """community_pooling docstring example (community_pooling.py:176-188) on two copies."""
    batch = pyg_min.Batch.from_data_list([_toy(), _toy()])
    cluster = torch.tensor([0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3, 3])
    out = pooling.community_pooling(cluster, batch)
    assert out.x.view(-1).tolist() == [2., 5., 2., 5.]
    assert out.edge_index.numel() == 0                      # every edge became a self loop
    assert out.batch.tolist() == [0, 0, 1, 1]
    assert out.pos.shape == (4, 3)
    # pairs {0,1} {2,3} {4,5}: edges 0-1 and 4-5 become self loops, 1-2 and 3-4 survive as c0-c1, c1-c2
    cluster = torch.tensor([0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5])
    out = pooling.community_pooling(cluster, pyg_min.Batch.from_data_list([_toy(), _toy()]))
    assert out.edge_index.tolist() == [[0, 1, 1, 2, 3, 4, 4, 5], [1, 0, 2, 1, 4, 3, 5, 4]]   # sorted by (row, col)
    assert out.x.view(-1).tolist() == [1., 3., 5., 1., 3., 5.]


def test_pooling_requires_pos_like_reference():
    d = _toy()
    d.pos = None
    with pytest.raises(UnboundLocalError):                   # community_pooling.py:226
        pooling.community_pooling(torch.tensor([0, 0, 0, 1, 1, 1]), pyg_min.Batch.from_data_list([d]))


def test_get_preloaded_cluster_literal_equals_closed_form():
    g = torch.Generator().manual_seed(0)
    sizes = [7, 1, 30, 2, 11]
    batch = torch.cat([torch.full((n,), i, dtype=torch.long) for i, n in enumerate(sizes)])
    cluster = torch.cat([torch.randint(0, 5, (n,), generator=g) for n in sizes])
    a = pooling.get_preloaded_cluster(cluster.clone(), batch)
    b = pooling.get_preloaded_cluster_closed_form(cluster.clone(), batch)
    assert torch.equal(a, b)
    # B = 1: the loop body never runs
    assert torch.equal(pooling.get_preloaded_cluster(cluster[:7].clone(), batch[:7]), cluster[:7])


def test_scatter_max_published_semantics():
    src = torch.tensor([[1., 0.], [1., 0.], [float('nan'), 3.], [2., float('nan')], [0., 0.]], requires_grad=True)
    index = torch.tensor([0, 0, 0, 2, 2])
    out, arg = pyg_min.scatter_max(src, index, dim=0, dim_size=4)
    assert out.tolist() == [[1., 3.], [0., 0.], [2., 0.], [0., 0.]]        # NaN never wins, empty -> 0
    assert arg.tolist() == [[0, 2], [5, 5], [3, 4], [5, 5]]               # first occurrence; empty -> src.size(0)
    out.sum().backward()
    assert src.grad.tolist() == [[1., 0.], [0., 0.], [0., 1.], [1., 0.], [0., 1.]]   # routed to argmax only


def test_scatter_mean_and_sum_semantics():
    src = torch.tensor([[1.], [3.], [5.]])
    idx = torch.tensor([2, 2, 0])
    assert pyg_min.scatter_mean(src, idx, dim=0).view(-1).tolist() == [5., 0., 2.]
    out = torch.zeros(4, 1)
    assert pyg_min.scatter_sum(src, idx, dim=0, out=out).view(-1).tolist() == [5., 0., 4., 0.]


def test_consecutive_cluster_and_pool_edge():
    inv, perm = pyg_min.consecutive_cluster(torch.tensor([5, 9, 5, 2, 9]))
    assert inv.tolist() == [1, 2, 1, 0, 2]
    assert [int(torch.tensor([5, 9, 5, 2, 9])[p]) for p in perm] == [2, 5, 9]
    ei = torch.tensor([[0, 1, 2, 3, 3, 0], [2, 3, 0, 1, 0, 3]])
    ea = torch.tensor([[1.], [2.], [4.], [8.], [16.], [32.]])
    cl = torch.tensor([0, 0, 1, 1])
    pei, pea = pyg_min.pool_edge(cl, ei, ea)
    assert pei.tolist() == [[0, 1], [1, 0]]
    assert pea.view(-1).tolist() == [1. + 2. + 32., 4. + 8. + 16.]


def test_batch_collation_rule():
    a, b = _toy(), _toy()
    a.cluster0 = torch.tensor([0, 0, 1, 1, 2, 2])
    b.cluster0 = torch.tensor([0, 1, 1, 1, 0, 0])
    a.mol, b.mol = 'a', 'b'
    batch = pyg_min.Batch.from_data_list([a, b])
    assert batch.edge_index[:, 8:].tolist() == (a.edge_index + 6).tolist()      # 'index' keys offset + cat on last dim
    assert batch.cluster0.tolist() == [0, 0, 1, 1, 2, 2, 0, 1, 1, 1, 0, 0]      # NOT offset (hence get_preloaded_cluster)
    assert batch.batch.tolist() == [0] * 6 + [1] * 6
    assert batch.mol == ['a', 'b']


# ------------------------------------------------------------------ algebraic identities
def test_ginet_attention_is_dead_code():
    """softmax over a size-1 dim == 1 (ginet.py:62-66): z == A (X W^T), attention grads are zero tensors."""
    torch.manual_seed(0)
    layer = onets.GINetConvLayer(5, 4)
    x = torch.randn(9, 5)
    ei = torch.randint(0, 9, (2, 30))
    ea = torch.rand(30, 1)
    z = layer(x, ei, ea)
    A = torch.zeros(9, 9).index_put_((ei[0], ei[1]), torch.ones(30), accumulate=True)
    torch.testing.assert_close(z, A @ (x @ layer.fc.weight.t()), rtol=1e-5, atol=1e-5)
    z.sum().backward()
    assert layer.fc_attention.weight.grad is not None and float(layer.fc_attention.weight.grad.abs().max()) == 0.0
    assert layer.fc_edge_attr.weight.grad is not None and float(layer.fc_edge_attr.weight.grad.abs().max()) == 0.0


def test_sgat_factorisation():
    """out_i = s_i (x_i W_top) + m_i W_bot + b with s_i = sum a_e / d_i, m_i = sum a_e x_col / d_i."""
    torch.manual_seed(1)
    layer = onets.sGraphAttentionLayer(6, 4)
    n, E = 11, 40
    x = torch.randn(n, 6)
    ei = torch.randint(0, n, (2, E))
    ea = torch.rand(E, 1) + 0.5
    ref = layer(x, ei, ea)
    row, col = ei
    deg = torch.bincount(row, minlength=n).clamp(min=1).float()
    s = torch.zeros(n).index_add_(0, row, ea.view(-1)) / deg
    m = torch.zeros(n, 6).index_add_(0, row, x[col] * ea) / deg.view(-1, 1)
    W = layer.weight
    out = s.view(-1, 1) * (x @ W[:6]) + m @ W[6:] + layer.bias
    torch.testing.assert_close(out, ref, rtol=1e-5, atol=1e-5)


def test_fout_literal_loop_equals_vectorised_and_nan_on_isolated():
    torch.manual_seed(2)
    layer = onets.FoutLayer(5, 3)
    n = 8
    x = torch.randn(n, 5)
    ei = torch.tensor([[0, 1, 2, 3, 4, 5, 6, 0], [1, 0, 3, 2, 5, 4, 0, 6]])     # node 7 has no neighbour
    onets.LITERAL = True
    a = layer(x, ei)
    onets.LITERAL = False
    b = layer(x, ei)
    onets.LITERAL = True
    assert torch.isnan(a[7]).all() and torch.isnan(b[7]).all()                   # foutnet.py:73 mean of empty
    torch.testing.assert_close(a[:7], b[:7], rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize('net', ['GINet', 'sGAT', 'FoutNet'])
def test_nets_literal_equals_vectorised_on_fixture(net):
    from deeprank_gnn_b200.DataSet import HDF5DataSet
    ds = HDF5DataSet(database=FIXTURE, node_feature=['type', 'polarity', 'bsa'], target='irmsd')
    graphs = [ds.get(i) for i in range(4)]
    cls = {'GINet': onets.GINet, 'sGAT': onets.sGAT, 'FoutNet': onets.FoutNet}[net]
    torch.manual_seed(0)
    model = cls(3, 1, 1).eval()
    onets.LITERAL = True
    a = model(to_oracle_batch(graphs))
    onets.LITERAL = False
    b = model(to_oracle_batch(graphs))
    onets.LITERAL = True
    assert a.shape == (4, 1)
    torch.testing.assert_close(a, b, rtol=1e-5, atol=1e-5)


# ------------------------------------------------------------------ fixture facts
FIX_ORDER = ['1ATN_10w', '1ATN_1w', '1ATN_2w', '1ATN_3w', '1ATN_4w', '1ATN_5w', '1ATN_6w', '1ATN_7w', '1ATN_8w',
             '1ATN_9w']
FIX_N = {'1w': 132, '2w': 132, '3w': 137, '4w': 95, '5w': 129, '6w': 118, '7w': 114, '8w': 106, '9w': 124, '10w': 108}
FIX_E = {'1w': 374, '2w': 330, '3w': 386, '4w': 201, '5w': 331, '6w': 298, '7w': 278, '8w': 253, '9w': 300, '10w': 252}
FIX_K0 = {'1w': 31, '2w': 33, '3w': 36, '4w': 23, '5w': 34, '6w': 28, '7w': 30, '8w': 25, '9w': 28, '10w': 29}
FIX_K1 = {'1w': 10, '2w': 9, '3w': 10, '4w': 9, '5w': 9, '6w': 11, '7w': 14, '8w': 8, '9w': 7, '10w': 11}


def test_fixture_structure_facts():
    from deeprank_gnn_b200.DataSet import HDF5DataSet
    ds = HDF5DataSet(database=FIXTURE, node_feature=['type', 'polarity', 'bsa'], target='irmsd')
    assert [m for _, m in ds.index_complexes] == FIX_ORDER
    graphs = [ds.get(i) for i in range(ds.len())]
    for g in graphs:
        k = g.mol.split('_')[1]
        assert g.x.shape == (FIX_N[k], 3)
        assert g.edge_index.shape == (2, 2 * FIX_E[k])
        assert g.cluster0.unique().numel() == FIX_K0[k] == g.cluster1.numel()
        assert g.cluster1.unique().numel() == FIX_K1[k]
        row, col = g.edge_index
        assert torch.equal(row[:FIX_E[k]], col[FIX_E[k]:]) and torch.equal(col[:FIX_E[k]], row[FIX_E[k]:])
        assert (torch.bincount(row, minlength=FIX_N[k]) > 0).all()          # no isolated node
        assert 0.0 < float(g.edge_attr.min()) and float(g.edge_attr.max()) < 2.0       # tanh(-d/2+2)+1
    st = oracle_structure(graphs)
    assert st['K0'] == sum(FIX_K0.values()) and st['K1'] == sum(FIX_K1.values())
    assert 4.5 < 2 * sum(FIX_E.values()) / st['E1'] < 5.3                      # dedupe ratio 4.7-5.2x
    ei1 = st['edge_index1']
    key = ei1[0] * st['K0'] + ei1[1]
    assert bool((key[1:] > key[:-1]).all())                                    # sorted by (row, col), unique
    assert bool((ei1[0] != ei1[1]).all())
    assert torch.equal(torch.sort(key)[0], torch.sort(ei1[1] * st['K0'] + ei1[0])[0])   # stays symmetric


def test_checkpoint_schema_matches_oracle_state_dict():
    """The 13 shipped GINet checkpoints pin parameter names / shapes (fixture generated by
    tests/golden/make_checkpoint_schema.py)."""
    with open(os.path.join(ROOT, 'tests', 'golden', 'checkpoint_schema.json')) as f:
        schema = json.load(f)
    assert len(schema) == 13
    for name, entry in schema.items():
        F = entry['model']['conv1.fc.weight'][1]
        out = entry['model']['fc2.weight'][0]
        model = onets.GINet(F, out, 1)
        sd = {k: list(v.shape) for k, v in model.state_dict().items()}
        assert sd == entry['model'], name
        assert entry['optimizer_state_entries'] in (0, 16)
    from deeprank_gnn_b200.engine import NetSpec
    spec = NetSpec('GINet', 48, 1, 1)
    assert sorted(n for n, _s, _l in spec.param_shapes()) == sorted(schema[next(iter(schema))]['model'].keys())
