"""Scoring with the SHIPPED pretrained weights: the reference GINet loaded with
``paper_pretrained_models/scoring_of_docking_models/fold6_treg_yfnat_b128_e20_lr0.001_4.pt`` scored 512 seeded
F = 48 graphs in the development container (``tests/golden/make_scoring_vectors.py``, the unmodified reference
``ginet.py``); here the oracle (CPU) and the CUDA scoring path (``Engine.train_batches(train=False)``, batches of
512 and of 128 = the checkpoint's batch size) must reproduce those predictions to absolute 1e-4."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'scoring_fold6.npz')


def _gold():
    z = np.load(GOLD, allow_pickle=False)
    sd = {k[len('model/'):]: torch.from_numpy(z[k]) for k in z.files if k.startswith('model/')}
    nodes, edges, feat = [int(v) for v in z['gen']]
    return z, sd, dict(nodes=nodes, edges=edges, feat=feat)


def _graphs(z, gen):
    from deeprank_gnn_b200 import synthetic
    graphs = synthetic.make_graphs(gen, count=int(z['count']), seed=int(z['seed']), internal=False)
    chk = float(sum(g.x.double().sum() for g in graphs))
    assert abs(chk - float(z['x_checksum'])) < 1e-6 * max(1.0, abs(chk)), 'the seeded generator drifted'
    return graphs


def test_oracle_reproduces_reference_scores_with_shipped_weights():
    from helpers import to_oracle_batch
    from oracle import nets as onets
    z, sd, gen = _gold()
    assert tuple(sd['conv1.fc.weight'].shape) == (16, 48) and str(z['target']) == 'fnat'
    graphs = _graphs(z, gen)[:128]
    model = onets.GINet(48, 1, 1).eval()
    model.load_state_dict(sd)
    with torch.no_grad():
        pred = model(to_oracle_batch(graphs)).reshape(-1)
    assert float((pred - torch.from_numpy(z['pred128'][:128])).abs().max()) < 1e-5


@pytest.mark.gpu
@pytest.mark.parametrize('batch', [512, 128])
def test_cuda_scoring_path_reproduces_reference_scores_with_shipped_weights(lib, batch):
    from deeprank_gnn_b200.data import Batch, PackedBatch
    from deeprank_gnn_b200.engine import Engine
    z, sd, gen = _gold()
    graphs = _graphs(z, gen)
    eng = Engine('GINet', 48, 1, 1, device='cuda:0', graph=True).eval()
    eng.load_state_dict(sd)
    packed = [PackedBatch.from_batch(Batch.from_data_list(graphs[i:i + batch]), idx16=True, edge_attr=False)
              for i in range(0, len(graphs), batch)]
    w0 = eng.params.data.clone()
    # one pass through the public scoring call (NeuralNet.test / eval use it): repeat the batches so that the C feeder
    # loop (more batches than pipeline slots) is exercised too
    reps = 2 if batch == 128 else 6
    _losses, preds = eng.train_batches(packed * reps, train=False)
    eng.validate()
    ref = torch.from_numpy(z['pred512'])
    n_b = len(packed)
    for rep in range(reps):
        got = torch.cat([p.reshape(-1) for p in preds[rep * n_b:(rep + 1) * n_b]])
        err = float((got - ref).abs().max())
        assert err <= 1e-4, 'batch %d pass %d: max|diff| = %.3e' % (batch, rep, err)
    assert torch.equal(eng.params.data, w0) and float(eng.step_dev[0]) == 0.0      # scoring leaves the weights alone
