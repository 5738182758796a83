"""Generate ``tests/golden/scoring_fold6.npz``: the REFERENCE GINet with the SHIPPED pretrained weights scoring
seeded synthetic interface graphs.

TEST INFRASTRUCTURE.  Run in the development container only (needs ``/root/reference``):

    python tests/golden/make_scoring_vectors.py

* the model is the unmodified ``/root/reference/deeprank_gnn/ginet.py::GINet`` (third-party modules shimmed
  exactly as in ``make_reference_vectors.py``) built the way ``NeuralNet.load_params`` /
  ``load_pretrained_model`` do (``NeuralNet.py:93-120, 794-823``): ``GINet(48, 1, 1)`` +
  ``load_state_dict(checkpoint['model'])`` of
  ``paper_pretrained_models/scoring_of_docking_models/fold6_treg_yfnat_b128_e20_lr0.001_4.pt``
  (7 node features = 48 columns after the one-hot encoding of ``ResidueGraph.py:239-244``);
* inputs: 512 graphs of the cfg2 shape with F = 48 from the product's seeded generator (no matching HDF5
  ships with the checkpoint - SURVEY 8c-3), scored in ``model.eval()`` as one batch of 512 and as four
  batches of 128 (the checkpoint's ``batch_size``);
* stored: the checkpoint's ``model`` tensors (45 KB - ``/root/reference`` does not exist on the GPU box), its
  hyper-parameters, the generator arguments and the reference predictions.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import make_reference_vectors as mrv  # noqa: E402
from oracle import pyg_min  # noqa: E402

CKPT = os.path.join(mrv.REF, 'paper_pretrained_models', 'scoring_of_docking_models', 'fold6_treg_yfnat_b128_e20_lr0.001_4.pt')
OUT = os.path.join(HERE, 'scoring_fold6.npz')
GEN = dict(nodes=200, edges=1000, feat=48)
COUNT, SEED = 512, 2024


def main():
    mrv.install_shims()
    from deeprank_gnn.ginet import GINet
    import deeprank_gnn.ginet as ref_ginet
    assert ref_ginet.__file__.startswith(mrv.REF)
    ck = torch.load(CKPT, map_location='cpu', weights_only=False)
    model = GINet(48, 1, 1)
    model.load_state_dict(ck['model'])
    model.eval()
    from deeprank_gnn_b200 import synthetic
    graphs = synthetic.make_graphs(GEN, count=COUNT, seed=SEED, internal=False)

    def batch(gs):
        return pyg_min.Batch.from_data_list(
            [pyg_min.Data(**{k: (g[k].clone() if torch.is_tensor(g[k]) else g[k]) for k in g.keys}) for g in gs])
    with torch.no_grad():
        pred512 = model(batch(graphs)).reshape(-1)
        pred128 = torch.cat([model(batch(graphs[i:i + 128])).reshape(-1) for i in range(0, COUNT, 128)])
    assert float((pred512 - pred128).abs().max()) < 1e-5
    out = {'pred512': pred512.numpy(), 'pred128': pred128.numpy(), 'count': COUNT, 'seed': SEED,
           'gen': np.array([GEN['nodes'], GEN['edges'], GEN['feat']]),
           'x_checksum': float(sum(g.x.double().sum() for g in graphs)),
           'node_features': np.array(ck['node']), 'target': ck['target'], 'task': ck['task'], 'batch_size': ck['batch_size'],
           'lr': ck['lr']}
    for k, v in ck['model'].items():
        out['model/' + k] = v.numpy()
    np.savez_compressed(OUT, **out)
    print('wrote %s (%.1f KB): pred range [%.4f, %.4f], mean %.4f' % (OUT, os.path.getsize(OUT) / 1024.0, float(pred512.min()),
                                                                       float(pred512.max()), float(pred512.mean())))


if __name__ == '__main__':
    main()
