"""Generate ``tests/golden/reference_vectors.npz`` by running the REFERENCE's own Python files.

TEST INFRASTRUCTURE.  Run in the development container only (needs ``/root/reference``):

    python tests/golden/make_reference_vectors.py

What runs here is the unmodified reference source, imported from where it lies
(``/root/reference/deeprank_gnn/{ginet,sGAT,foutnet,community_pooling,DataSet,NeuralNet,Metrics}.py``):
its layers, its ``GINet/sGAT/FoutNet.forward``, ``get_preloaded_cluster``, ``community_pooling``,
``HDF5DataSet.load_one_graph`` and ``NeuralNet.train/_epoch/eval`` - NOT a restatement of them.
What is NOT the real thing are the third-party modules those files import, none of which is
installed in this image (no network): ``torch_scatter``, ``torch_geometric``, ``h5py``,
``community``, ``markov_clustering``.  They are injected into ``sys.modules`` as thin shims:

* ``torch_scatter`` / ``torch_geometric``  -> ``oracle.pyg_min`` (published-behaviour restatement);
* ``h5py``                                  -> the bundled read-only parser ``hdf5min`` for ``'r'``,
                                              an in-memory sink for ``'w'`` (epoch exports are discarded);
* ``community`` / ``markov_clustering``     -> empty modules (``PreCluster`` is replaced by a no-op:
                                              the fixture already stores its MCL clusters, which the
                                              reference would recompute with the same method).

So the vectors pin the oracle's restatement of the REFERENCE FILES (and through it the CUDA path)
against the reference code itself; the third-party primitives stay restated.
The output is consumed by ``tests/test_reference_vectors.py`` (CPU: oracle; GPU: CUDA path).
"""
import copy
import io
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get('DRGNN_REFERENCE', '/root/reference')
FIXTURE = os.path.join(HERE, '1ATN_residue.hdf5')
OUT = os.path.join(HERE, 'reference_vectors.npz')
sys.path.insert(0, ROOT)

from oracle import pyg_min  # noqa: E402
from deeprank_gnn_b200 import hdf5min  # noqa: E402  (file parser only; no compute)


# --------------------------------------------------------------------------- shims
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def _unavailable(name):
    def f(*a, **k):
        raise RuntimeError('%s is not available in the shim (not on the hot path)' % name)
    return f


class _Dataset(object):
    """torch_geometric.data.dataset.Dataset: only what HDF5DataSet / DataLoader use."""

    def __init__(self, root=None, transform=None, pre_transform=None):
        self.root, self.transform, self.pre_transform = root, transform, pre_transform

    def __len__(self):
        return self.len()

    def __getitem__(self, idx):
        d = self.get(idx)
        return d if self.transform is None else self.transform(d)


class _DataLoader(object):
    """torch_geometric.data.DataLoader: sequential or torch-RNG-shuffled mini-batches collated by
    Batch.from_data_list (NeuralNet.py:105,153,158)."""

    def __init__(self, dataset, batch_size=1, shuffle=False, **kw):
        self.dataset, self.batch_size, self.shuffle = dataset, batch_size, shuffle

    def __iter__(self):
        n = len(self.dataset)
        order = torch.randperm(n).tolist() if self.shuffle else list(range(n))
        for s in range(0, n, self.batch_size):
            yield pyg_min.Batch.from_data_list([self.dataset[i] for i in order[s:s + self.batch_size]])

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size


class _Sink(object):
    """h5py.File(..., 'w'): swallows the epoch export (NeuralNet.py:827-872)."""

    def __init__(self):
        self.attrs = {}

    def create_group(self, name):
        return _Sink()

    require_group = create_group

    def create_dataset(self, name, data=None, dtype=None):
        return None

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _h5_file(path, mode='r'):
    if mode == 'r':
        return hdf5min.File(path, 'r')
    return _Sink()


def install_shims():
    _mod('torch_scatter', scatter_sum=pyg_min.scatter_sum, scatter_add=pyg_min.scatter_sum,
         scatter_mean=pyg_min.scatter_mean, scatter_max=pyg_min.scatter_max)
    tg = _mod('torch_geometric')
    tg.utils = _mod('torch_geometric.utils', remove_self_loops=pyg_min.remove_self_loops,
                    add_self_loops=_unavailable('add_self_loops'), softmax=_unavailable('softmax'))
    tg.nn = _mod('torch_geometric.nn', max_pool_x=pyg_min.max_pool_x)
    tg.nn.inits = _mod('torch_geometric.nn.inits', uniform=pyg_min.uniform)
    tg.nn.pool = _mod('torch_geometric.nn.pool')
    tg.nn.pool.pool = _mod('torch_geometric.nn.pool.pool', pool_edge=pyg_min.pool_edge,
                           pool_batch=pyg_min.pool_batch, pool_pos=_unavailable('pool_pos'))
    tg.nn.pool.consecutive = _mod('torch_geometric.nn.pool.consecutive',
                                  consecutive_cluster=pyg_min.consecutive_cluster)
    tg.data = _mod('torch_geometric.data', Data=pyg_min.Data, Batch=pyg_min.Batch, DataLoader=_DataLoader)
    tg.data.dataset = _mod('torch_geometric.data.dataset', Dataset=_Dataset)
    tg.data.data = _mod('torch_geometric.data.data', Data=pyg_min.Data)
    _mod('h5py', File=_h5_file, special_dtype=lambda **k: None)
    _mod('community')
    _mod('markov_clustering')
    # the reference package, WITHOUT its __init__ (which pulls the graph-generation stack)
    pkg = types.ModuleType('deeprank_gnn')
    pkg.__path__ = [os.path.join(REF, 'deeprank_gnn')]
    sys.modules['deeprank_gnn'] = pkg


# --------------------------------------------------------------------------- helpers
OUTD = {}


def put(key, val):
    if torch.is_tensor(val):
        val = val.detach().cpu().numpy()
    OUTD[key] = np.array(val, copy=True)              # a state_dict entry aliases the live parameter


def put_state(prefix, sd):
    for k, v in sd.items():
        put('%s/%s' % (prefix, k), v)


def quiet(fn, *a, **k):
    so = sys.stdout
    sys.stdout = io.StringIO()
    try:
        return fn(*a, **k)
    finally:
        sys.stdout = so


def one_step(tag, Net, make_batch, nfeat, out_shape=1, seed=0):
    """Reference net: seeded init, eval forward, MSE / CE loss, backward, one Adam(lr=0.01) step
    (NeuralNet.py:490-503 without the loader)."""
    torch.manual_seed(seed)
    model = Net(nfeat, out_shape, 1)
    model.eval()                                      # no dropout: deterministic on any device
    put_state(tag + '/state0', model.state_dict())
    batch = make_batch()
    y = batch.y.clone()
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    opt.zero_grad()
    pred = model(batch)
    if out_shape == 1:
        loss = torch.nn.MSELoss()(pred.reshape(-1), y)
    else:
        tgt = (y > y.median()).long()
        put(tag + '/class_target', tgt)
        loss = torch.nn.CrossEntropyLoss(weight=torch.tensor([0.3, 0.7]), reduction='mean')(pred, tgt)
    loss.backward()
    put(tag + '/pred', pred)
    put(tag + '/loss', loss)
    for n, p in model.named_parameters():
        put('%s/grad/%s' % (tag, n), p.grad)
    opt.step()
    put_state(tag + '/state1', model.state_dict())


def main():
    install_shims()
    from deeprank_gnn import community_pooling as ref_cp
    from deeprank_gnn import DataSet as ref_ds
    from deeprank_gnn import NeuralNet as ref_nn
    from deeprank_gnn.ginet import GINet
    from deeprank_gnn.sGAT import sGAT
    from deeprank_gnn.foutnet import FoutNet
    for m in (ref_cp, ref_ds, ref_nn):
        assert m.__file__.startswith(REF), m.__file__
    ref_ds.PreCluster = lambda dataset, method: None
    ref_nn.PreCluster = lambda dataset, method: None
    # Metrics.py:200-201 passes ``squared=`` which the installed scikit-learn no longer accepts
    # (reporting only, not on the path): accept and honour it.
    if not hasattr(np, 'string_'):
        np.string_ = np.bytes_                        # NeuralNet.py:862 (export of mol names), NumPy >= 2
    from sklearn import metrics as _skm
    _mse = _skm.mean_squared_error
    _skm.mean_squared_error = lambda y, p, squared=True, **k: (_mse(y, p, **k) if squared else _mse(y, p, **k) ** 0.5)

    # ---- (1) a14: records built by the reference loader from the shipped fixture (cfg1 features)
    ds = quiet(ref_ds.HDF5DataSet, root='./', database=FIXTURE, node_feature=['type', 'polarity', 'bsa'],
               edge_feature=['dist'], target='irmsd', tqdm=False)
    put('records/count', len(ds))
    mols = []
    for i in range(len(ds)):
        d = ds.get(i)
        mols.append(d.mol)
        for k in ('x', 'edge_index', 'edge_attr', 'internal_edge_index', 'internal_edge_attr', 'y', 'pos',
                  'cluster0', 'cluster1'):
            put('records/%d/%s' % (i, k), d[k])
    put('records/mol', np.array(mols))

    def fixture_batch(count=8, ds=ds):
        return pyg_min.Batch.from_data_list([ds.get(i) for i in range(count)])

    # ---- (2) a8/a9/a10: reference get_preloaded_cluster + community_pooling on the cfg1 batch
    b = fixture_batch()
    c0 = ref_cp.get_preloaded_cluster(b.cluster0, b.batch)
    put('pool/cluster0_offset', c0)
    pooled = ref_cp.community_pooling(c0, b)
    for k in ('x', 'edge_index', 'edge_attr', 'internal_edge_index', 'internal_edge_attr', 'pos', 'batch'):
        put('pool/' + k, pooled[k])
    c1 = ref_cp.get_preloaded_cluster(pooled.cluster1, pooled.batch)
    put('pool/cluster1_offset', c1)
    x2, batch2 = sys.modules['torch_geometric.nn'].max_pool_x(c1, pooled.x, pooled.batch)
    put('pool/x2', x2)
    put('pool/batch2', batch2)

    # ---- (3) a1-a7, a11-a13: the three reference nets, one training step on the cfg1 batch
    for name, Net in (('GINet', GINet), ('sGAT', sGAT), ('FoutNet', FoutNet)):
        one_step('cfg1/' + name, Net, fixture_batch, 3)
    one_step('cfg1_class/GINet', GINet, fixture_batch, 3, out_shape=2)

    # ---- (4) the test-suite feature list of the reference (tests/test_nn.py:11-13): F = 28
    feats = ['type', 'polarity', 'bsa', 'depth', 'hse', 'ic', 'pssm']
    ds28 = quiet(ref_ds.HDF5DataSet, root='./', database=FIXTURE, node_feature=feats, edge_feature=['dist'],
                 target='irmsd', tqdm=False)
    put('f28/x0', ds28.get(0).x)
    for name, Net in (('GINet', GINet), ('sGAT', sGAT), ('FoutNet', FoutNet)):
        one_step('f28/' + name, Net, lambda: fixture_batch(10, ds28), 28)

    # ---- (5) synthetic cfg2-shaped graphs (product generator, seeded): 4 graphs, F = 32
    from deeprank_gnn_b200 import synthetic
    graphs = synthetic.make_graphs('cfg2', count=4, seed=123)

    def syn_batch():
        return pyg_min.Batch.from_data_list(
            [pyg_min.Data(**{k: (g[k].clone() if torch.is_tensor(g[k]) else g[k]) for k in g.keys}) for g in graphs])
    sb = syn_batch()
    put('cfg2/x_checksum', sb.x.double().sum())
    put('cfg2/edge_checksum', (sb.edge_index.double() * torch.arange(1, sb.edge_index.size(1) + 1)).sum())
    for name, Net in (('GINet', GINet), ('sGAT', sGAT), ('FoutNet', FoutNet)):
        one_step('cfg2/' + name, Net, syn_batch, 32)

    # ---- (6) a13 through the reference driver: NeuralNet(...).train(nepoch=3) on the fixture,
    #          sGAT / FoutNet (no dropout => RNG-free trajectory), batch 4, no shuffle
    for name, Net in (('sGAT', sGAT), ('FoutNet', FoutNet)):
        torch.manual_seed(7)
        np.random.seed(7)
        nn_ = quiet(ref_nn.NeuralNet, FIXTURE, Net, node_feature=['type', 'polarity', 'bsa'], edge_feature=['dist'],
                    target='irmsd', lr=0.01, batch_size=4, percent=[1.0, 0.0], shuffle=False, outdir='/tmp')
        put('train/%s/order' % name, np.array([m for _, m in nn_.train_loader.dataset.index_complexes]))
        put_state('train/%s/state0' % name, copy.deepcopy(nn_.model.state_dict()))
        quiet(nn_.train, nepoch=3, validate=False, save_model='none', hdf5='drgnn_golden_sink.hdf5')
        put('train/%s/epoch_loss' % name, np.array(nn_.train_loss, dtype=np.float64))
        put('train/%s/last_outputs' % name, np.array(nn_.train_out, dtype=np.float64))
        put('train/%s/last_targets' % name, np.array(nn_.train_y, dtype=np.float64))
        put_state('train/%s/state3' % name, nn_.model.state_dict())

    np.savez_compressed(OUT, **OUTD)
    print('wrote %s: %d arrays, %.1f KB' % (OUT, len(OUTD), os.path.getsize(OUT) / 1024.0))


if __name__ == '__main__':
    main()
