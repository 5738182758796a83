"""Drop-in training / scoring driver: ``NeuralNet(database, Net, ...)`` with the reference's
constructor, ``train`` / ``test`` / ``eval`` / ``save_model`` / ``load_params`` methods and
checkpoint schema (``deeprank_gnn/NeuralNet.py:18-872``).

Two execution paths behind the same API:

* **fused** (default when ``Net`` is one of this package's GINet / sGAT / FoutNet): batches are
  packed once into pinned host blocks and an epoch is ``Engine.train_batches`` - one H2D copy,
  one fused kernel sequence and one asynchronous read-back per batch, a single host
  synchronisation per epoch (the reference synchronises several times per batch,
  NeuralNet.py:501-523).  Parameters are mirrored into ``self.model`` so ``state_dict`` /
  ``save_model`` behave as in the reference.
* **autograd** (any user-defined ``Net``, or ``fused=False``): the reference's loop
  (zero_grad, forward, loss, backward, Adam step) over the ``nn.Module``.

Plots and the PyQt/HDF5 explorer are out of scope; epoch exports go to HDF5 when h5py is
importable and to ``.npz`` files otherwise.
"""
import os
from time import time

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from .DataSet import DivideDataSet, HDF5DataSet, PreCluster
from .Metrics import Metrics
from .data import DataLoader, PackedBatch

try:                                    # pragma: no cover
    import h5py as _h5py
except Exception:                       # noqa
    _h5py = None


class _EngineOptimizer(object):
    """``optimizer.state_dict()`` / ``load_state_dict()`` facade over the engine's flat Adam state."""

    def __init__(self, engine):
        self.engine = engine

    def state_dict(self):
        return self.engine.optimizer_state_dict()

    def load_state_dict(self, sd):
        self.engine.load_optimizer_state_dict(sd)

    def zero_grad(self):
        pass


class _EpochWriter(object):
    """HDF5 (h5py) or .npz sink with the reference's epoch layout (NeuralNet.py:827-872)."""

    def __init__(self, fname):
        self.fname = fname
        self.h5 = _h5py.File(fname, 'w') if _h5py is not None else None
        self.npz = {}

    def export(self, epoch, data, attrs):
        name = 'epoch_%04d' % epoch
        if self.h5 is not None:
            grp = self.h5.create_group(name)
            for k, v in attrs.items():
                grp.attrs[k] = v
            for pass_type, pass_data in data.items():
                sg = grp.create_group(pass_type)
                for key, val in pass_data.items():
                    if key == 'mol':
                        sg.create_dataset(key, data=np.array(val, dtype=object), dtype=_h5py.special_dtype(vlen=str))
                    else:
                        sg.create_dataset(key, data=val)
        else:
            for pass_type, pass_data in data.items():
                for key, val in pass_data.items():
                    self.npz['%s/%s/%s' % (name, pass_type, key)] = np.array(val)

    def close(self):
        if self.h5 is not None:
            self.h5.close()
        else:
            np.savez(os.path.splitext(self.fname)[0] + '.npz', **self.npz)


class NeuralNet(object):
    def __init__(self, database, Net, node_feature=['type', 'polarity', 'bsa'], edge_feature=['dist'], target='irmsd',
                 lr=0.01, batch_size=32, percent=[1.0, 0.0], database_eval=None, index=None, class_weights=None,
                 task=None, classes=[0, 1], threshold=None, pretrained_model=None, shuffle=True, outdir='./',
                 cluster_nodes='mcl', transform_sigmoid=False, fused=True, device=None, verbose=True, cache=None):
        """``cache`` (fused path only): a directory; the packed feeder records of every loader are written there
        once (``data.PackedCache``: one memory-mapped file per loader, page-locked with ``cudaHostRegister``) and
        every later epoch feeds the GPU straight from the mapping - no HDF5 access, no Python collation
        (``DataSet.py:231-366`` does both per graph per epoch).  The mini-batches are then FROZEN as composed in
        the first epoch; ``shuffle=True`` reshuffles their order, not their members."""
        self.fused = fused
        self.cache = cache
        self._records = {}
        self.verbose = verbose
        self._device_arg = device
        if pretrained_model is None:
            self.node_feature, self.edge_feature, self.target, self.lr = node_feature, edge_feature, target, lr
            self.batch_size, self.percent, self.index, self.class_weights = batch_size, percent, index, class_weights
            self.task, self.classes, self.threshold, self.shuffle = task, classes, threshold, shuffle
            self.outdir, self.cluster_nodes, self.transform_sigmoid = outdir, cluster_nodes, transform_sigmoid
            self.pretrained_model = None
            if self.task is None:
                if self.target in ('irmsd', 'lrmsd', 'fnat', 'dockQ'):
                    self.task = 'reg'
                elif self.target in ('bin_class', 'capri_classes'):
                    self.task = 'class'
                else:
                    raise ValueError("User target detected -> The task argument is required ('class' or 'reg').")
            if self.threshold is None:
                self.threshold = self.classes[1] if self.task == 'class' else 0.3
            self.load_model(database, Net, database_eval)
        else:
            self.pretrained_model = pretrained_model
            self.load_params(pretrained_model)
            self.outdir = outdir
            self.load_pretrained_model(database, Net)

    def _say(self, *a):
        if self.verbose:
            print(*a)

    # ------------------------------------------------------------------ construction
    def _dataset(self, database, index=None):
        ds = HDF5DataSet(root='./', database=database, index=index, node_feature=self.node_feature,
                         edge_feature=self.edge_feature, target=self.target, clustering_method=self.cluster_nodes)
        if self.cluster_nodes is not None:
            if self.cluster_nodes in ('mcl', 'louvain'):
                PreCluster(ds, method=self.cluster_nodes)
            else:
                raise ValueError("Invalid node clustering method. Please set cluster_nodes to 'mcl', 'louvain' or None.")
        return ds

    def load_pretrained_model(self, database, Net):
        test_dataset = self._dataset(database)
        self.test_loader = DataLoader(test_dataset, batch_size=self.batch_size if self.fused else 1)
        self._say('Test set loaded')
        self.put_model_to_device(test_dataset, Net)
        self.set_loss()
        self._make_optimizer()
        self.optimizer.load_state_dict(self.opt_loaded_state_dict)
        self.model.load_state_dict(self.model_load_state_dict)
        if self.engine is not None:
            self.engine.load_state_dict(self.model_load_state_dict)

    def load_model(self, database, Net, database_eval):
        dataset = self._dataset(database, self.index)
        train_dataset, valid_dataset = DivideDataSet(dataset, percent=self.percent)
        self.train_loader = DataLoader(train_dataset, batch_size=self.batch_size, shuffle=self.shuffle)
        self._say('Training set loaded')
        if self.percent[1] > 0.0:
            self.valid_loader = DataLoader(valid_dataset, batch_size=self.batch_size, shuffle=self.shuffle)
            self._say('Evaluation set loaded')
        if database_eval is not None:
            valid_dataset = self._dataset(database_eval, self.index)
            self.valid_loader = DataLoader(valid_dataset, batch_size=self.batch_size, shuffle=self.shuffle)
            self._say('Independent validation set loaded !')
        self.put_model_to_device(dataset, Net)
        self.set_loss()
        self._make_optimizer()
        self.train_acc, self.train_loss, self.valid_acc, self.valid_loss = [], [], [], []

    def put_model_to_device(self, dataset, Net):
        if self._device_arg is not None:
            self.device = torch.device(self._device_arg)
        else:
            self.device = torch.device('cuda' if torch.cuda.is_available() else 'cpu')
        self._say('device set to :', self.device)
        self.num_edge_features = len(self.edge_feature)
        nfeat = dataset.get(0).num_features
        if self.task == 'reg':
            self.output_shape = 1
        else:
            self.classes_to_idx = {i: idx for idx, i in enumerate(self.classes)}
            self.idx_to_classes = {idx: i for idx, i in enumerate(self.classes)}
            self.output_shape = len(self.classes)
        try:
            self.model = Net(nfeat, self.output_shape, self.num_edge_features).to(self.device)
        except TypeError:
            raise ValueError('The loaded model does not accept output_shape = %d argument' % self.output_shape)
        self.engine = None
        kind = type(self.model).__name__
        from . import foutnet, ginet, sGAT
        builtin = isinstance(self.model, (ginet.GINet, sGAT.sGAT, foutnet.FoutNet))
        if self.fused and builtin and self.device.type == 'cuda':
            from .engine import Engine
            hidden = getattr(self.model, 'hidden', None) or (self.model.conv1.out_channels, self.model.conv2.out_channels)
            self.engine = Engine(kind, nfeat, self.output_shape, self.num_edge_features, hidden=hidden,
                                 device=self.device, task=self.task, transform_sigmoid=self.transform_sigmoid,
                                 lr=self.lr, dropout=getattr(self.model, 'dropout', None))
            self.engine.load_state_dict(self.model.state_dict())

    def _make_optimizer(self):
        if self.engine is not None:
            self.optimizer = _EngineOptimizer(self.engine)
        else:
            self.optimizer = torch.optim.Adam(self.model.parameters(), lr=self.lr)

    def set_loss(self):
        self.weights = None
        if self.task == 'reg':
            self.loss = nn.MSELoss()
        else:
            if self.class_weights is True:
                self.weights = self.compute_class_weights()
            elif self.class_weights is not None and self.class_weights is not False:
                self.weights = torch.as_tensor(self.class_weights, dtype=torch.float32)
            w = None if self.weights is None else self.weights.to(self.device)
            self.loss = nn.CrossEntropyLoss(weight=w, reduction='mean')
            if self.engine is not None:
                self.engine.class_weights = w

    def compute_class_weights(self):
        targets_all = []
        for batch in self.train_loader:
            targets_all.append(batch.y)
        targets_all = torch.cat(targets_all).reshape(-1).tolist()
        weights = torch.tensor([targets_all.count(i) for i in self.classes], dtype=torch.float32)
        weights = 1.0 / weights
        return weights / weights.sum()

    # ------------------------------------------------------------------ epochs
    def format_output(self, pred, target=None):
        if self.task == 'class':
            if target is not None:
                target = torch.tensor([self.classes_to_idx[int(x)] for x in target])
        elif self.transform_sigmoid is True:
            pred = torch.sigmoid(pred.reshape(-1))
        else:
            pred = pred.reshape(-1)
        return pred, target

    def _collect(self, data, out, raw, y, pred, targets, mols):
        """Append one batch of (host) predictions / targets in the reference's export format."""
        if self.task == 'class':
            prob = F.softmax(pred, dim=1)
            raw += prob.tolist()
            out += prob.argmax(dim=1).tolist()
        else:
            p = pred.reshape(-1)
            if self.transform_sigmoid is True:
                p = torch.sigmoid(p)
            raw += p.tolist()
            out += p.tolist()
        if targets is not None:
            y += targets
        data['mol'] += list(mols)

    def _finish(self, data, out, raw, y):
        if self.task == 'class':
            data['targets'] += [self.idx_to_classes[x] for x in y]
            data['outputs'] += [self.idx_to_classes[x] for x in out]
        else:
            data['targets'] += y
            data['outputs'] += out
        data['raw_outputs'] += raw
        return data

    def _cached_records(self, loader):
        """(packed records, targets, loss normalisers, mol names) of ``loader`` from the packed cache, built on
        first use."""
        import hashlib
        import json
        from .data import PackedCache
        key = id(loader)
        if key in self._records:
            return self._records[key]
        os.makedirs(self.cache, exist_ok=True)
        ds = loader.dataset
        ident = json.dumps([[str(f), str(m)] for f, m in getattr(ds, 'index_complexes', [])] +
                           [self.batch_size, list(self.node_feature), list(self.edge_feature), str(self.target),
                            self.task, self.engine.spec.kind])
        stem = os.path.join(self.cache, 'records_' + hashlib.sha1(ident.encode()).hexdigest()[:16])
        if not (os.path.exists(stem + '.pack') and os.path.exists(stem + '.json')):
            shuffle, loader.shuffle = loader.shuffle, False        # composition frozen in data-set order
            try:
                packed, inv, tg, mols = self._pack(list(loader))
            finally:
                loader.shuffle = shuffle
            PackedCache.build(stem + '.pack', packed)
            with open(stem + '.json', 'w') as f:
                json.dump({'inv': inv, 'targets': tg, 'mol': mols}, f)
        cache = PackedCache(stem + '.pack', register=True)
        with open(stem + '.json') as f:
            meta = json.load(f)
        rec = (cache, [cache[i] for i in range(len(cache))], meta['targets'], meta['inv'], meta['mol'])
        self._records[key] = rec
        return rec

    def _pack(self, batches):
        eng = self.engine
        classes = self.classes if self.task == 'class' else None
        packed, inv, tg, mols = [], [], [], []
        for b in batches:
            has_y = getattr(b, 'y', None) is not None
            # compact feeder records: uint16 graph-local edge ids; edge attributes only for the net that reads them
            packed.append(PackedBatch.from_batch(b, classes=classes if has_y else None, idx16=True,
                                                 edge_attr=eng.spec.kind == 'sgat'))
            mols.append(list(b.mol) if isinstance(b.mol, list) else [b.mol])
            if not has_y:
                tg.append(None)
                inv.append(1.0)
            elif self.task == 'class':
                idx = [self.classes_to_idx[int(t)] for t in b.y.reshape(-1).tolist()]
                tg.append(idx)
                w = self.weights
                inv.append(1.0 / (float(w[idx].sum()) if w is not None else len(idx)))
            else:
                tg.append(b.y.reshape(-1).tolist())
                inv.append(1.0 / b.num_graphs)
        return packed, inv, tg, mols

    def _fused_pass(self, loader, train):
        eng = self.engine
        if self.cache:
            _cache, packed, tg, inv, mols = self._cached_records(loader)
            order = torch.randperm(len(packed)).tolist() if (loader.shuffle and train) else list(range(len(packed)))
            packed, tg, inv, mols = ([packed[i] for i in order], [tg[i] for i in order], [inv[i] for i in order],
                                     [mols[i] for i in order])
            losses, preds = eng.train_batches(packed, inv_norms=inv, train=train)
            eng.validate()
            out, raw, y = [], [], []
            data = {'outputs': [], 'raw_outputs': [], 'targets': [], 'mol': []}
            loss_val = 0.0
            for mol, l, p, t in zip(mols, losses.tolist(), preds, tg):
                if t is not None:
                    loss_val += l
                self._collect(data, out, raw, y, p, t, mol)
            if train:
                self.model.load_state_dict(eng.state_dict())
            return out, y, loss_val, self._finish(data, out, raw, y)
        batches = list(loader)
        classes = self.classes if self.task == 'class' else None
        packed, inv, tg = [], [], []
        for b in batches:
            has_y = getattr(b, 'y', None) is not None
            # compact feeder records: uint16 graph-local edge ids; edge attributes only for the net that reads them
            packed.append(PackedBatch.from_batch(b, classes=classes if has_y else None, idx16=True,
                                                 edge_attr=eng.spec.kind == 'sgat'))
            if not has_y:
                tg.append(None)
                inv.append(1.0)
            elif self.task == 'class':
                idx = [self.classes_to_idx[int(t)] for t in b.y.reshape(-1).tolist()]
                tg.append(idx)
                w = self.weights
                inv.append(1.0 / (float(w[idx].sum()) if w is not None else len(idx)))
            else:
                tg.append(b.y.reshape(-1).tolist())
                inv.append(1.0 / b.num_graphs)
        losses, preds = eng.train_batches(packed, inv_norms=inv, train=train)
        eng.validate()
        out, raw, y = [], [], []
        data = {'outputs': [], 'raw_outputs': [], 'targets': [], 'mol': []}
        loss_val = 0.0
        for b, l, p, t in zip(batches, losses.tolist(), preds, tg):
            if t is not None:
                loss_val += l
            self._collect(data, out, raw, y, p, t, b.mol if isinstance(b.mol, list) else [b.mol])
        if train:
            self.model.load_state_dict(eng.state_dict())
        return out, y, loss_val, self._finish(data, out, raw, y)

    def eval(self, loader):
        self.model.eval()
        if self.engine is not None:
            return self._fused_pass(loader, train=False)
        loss_val, out, raw, y = 0, [], [], []
        data = {'outputs': [], 'raw_outputs': [], 'targets': [], 'mol': []}
        with torch.no_grad():
            for data_batch in loader:
                mols = data_batch.mol
                data_batch = data_batch.to(self.device)
                pred = self.model(data_batch)
                pred, tgt = self.format_output(pred, data_batch.y)
                t = None
                if tgt is not None:
                    tgt = tgt.to(self.device)
                    loss_val += self.loss(pred, tgt).item()
                    t = tgt.tolist()
                raw_pred = pred.detach().cpu()
                if self.task != 'class' and self.transform_sigmoid is True:
                    raw_pred = torch.logit(raw_pred)
                self._collect(data, out, raw, y, raw_pred, t, mols)
        return out, y, loss_val, self._finish(data, out, raw, y)

    def _epoch(self, epoch):
        if self.engine is not None:
            return self._fused_pass(self.train_loader, train=True)
        running_loss, out, raw, y = 0, [], [], []
        data = {'outputs': [], 'raw_outputs': [], 'targets': [], 'mol': []}
        for data_batch in self.train_loader:
            mols = data_batch.mol
            data_batch = data_batch.to(self.device)
            self.optimizer.zero_grad()
            pred = self.model(data_batch)
            pred, tgt = self.format_output(pred, data_batch.y)
            if tgt is None:
                raise ValueError('You must provide target values (y) for the training set')
            tgt = tgt.to(self.device)
            loss = self.loss(pred, tgt)
            running_loss += loss.detach().item()
            loss.backward()
            self.optimizer.step()
            raw_pred = pred.detach().cpu()
            if self.task != 'class' and self.transform_sigmoid is True:
                raw_pred = torch.logit(raw_pred)
            self._collect(data, out, raw, y, raw_pred, tgt.tolist(), mols)
        return out, y, running_loss, self._finish(data, out, raw, y)

    def train(self, nepoch=1, validate=False, save_model='last', hdf5='train_data.hdf5', save_epoch='intermediate',
              save_every=5):
        fname = self.update_name(hdf5, self.outdir)
        self.f5 = _EpochWriter(fname)
        attrs = {'task': self.task, 'target': str(self.target), 'batch_size': self.batch_size}
        try:
            self.nepoch = nepoch
            self.data = {}
            for epoch in range(1, nepoch + 1):
                self.model.train()
                if self.engine is not None:
                    self.engine.train()
                t0 = time()
                _out, _y, _loss, self.data['train'] = self._epoch(epoch)
                t = time() - t0
                self.train_loss.append(_loss)
                self.train_out, self.train_y = _out, _y
                _acc = self.get_metrics('train', self.threshold).accuracy
                self.train_acc.append(_acc)
                self.print_epoch_data('train', epoch, _loss, _acc, t)
                if validate is True:
                    t0 = time()
                    _out, _y, _val_loss, self.data['eval'] = self.eval(self.valid_loader)
                    t = time() - t0
                    self.valid_loss.append(_val_loss)
                    self.valid_out, self.valid_y = _out, _y
                    _val_acc = self.get_metrics('eval', self.threshold).accuracy
                    self.valid_acc.append(_val_acc)
                    self.print_epoch_data('valid', epoch, _val_loss, _val_acc, t)
                    best = min(self.valid_loss) == _val_loss
                else:
                    best = min(self.train_loss) == _loss
                if save_model == 'best' and best:
                    self.save_model(filename='t{}_y{}_b{}_e{}_lr{}_{}.pth.tar'.format(
                        self.task, self.target, str(self.batch_size), str(nepoch), str(self.lr), str(epoch)))
                if save_epoch == 'all' or epoch == nepoch or \
                        (save_epoch == 'intermediate' and epoch % save_every == 0):
                    self.f5.export(epoch, self.data, attrs)
            if save_model == 'last':
                self.save_model(filename='t{}_y{}_b{}_e{}_lr{}.pth.tar'.format(
                    self.task, self.target, str(self.batch_size), str(nepoch), str(self.lr)))
        finally:
            self.f5.close()

    def test(self, database_test=None, threshold=4, hdf5='test_data.hdf5'):
        fname = self.update_name(hdf5, self.outdir)
        self.f5 = _EpochWriter(fname)
        try:
            if database_test is not None:
                test_dataset = self._dataset(database_test)
                self.test_loader = DataLoader(test_dataset, batch_size=self.batch_size if self.fused else 1)
            elif not hasattr(self, 'test_loader'):
                raise ValueError('You need to upload a test dataset: model.test(test_dataset)')
            self.data = {}
            _out, _y, _test_loss, self.data['test'] = self.eval(self.test_loader)
            self.test_out = _out
            if len(_y) == 0:
                self.test_y, self.test_acc = None, None
            else:
                self.test_y = _y
                self.test_acc = self.get_metrics('test', threshold).accuracy
            self.test_loss = _test_loss
            self.f5.export(0, self.data, {'task': self.task, 'target': str(self.target),
                                          'batch_size': self.batch_size})
        finally:
            self.f5.close()

    def get_metrics(self, data='eval', threshold=4.0, binary=True):
        if self.task == 'class':
            # the reference indexes classes_to_idx[threshold] directly (NeuralNet.py:549) and raises for
            # test()'s default threshold of 4: fall back to the model's own threshold instead
            threshold = self.classes_to_idx[threshold if threshold in self.classes_to_idx else self.threshold]
        pred, y = {'eval': (getattr(self, 'valid_out', []), getattr(self, 'valid_y', [])),
                   'train': (getattr(self, 'train_out', []), getattr(self, 'train_y', [])),
                   'test': (getattr(self, 'test_out', []), getattr(self, 'test_y', []))}[data]
        return Metrics(pred, y, self.target, threshold, binary)

    def print_epoch_data(self, stage, epoch, loss, acc, time):
        acc_str = 'None' if acc is None else '%1.4e' % acc
        self._say('Epoch [%04d] : %s loss %e | accuracy %s | time %1.2e sec.' % (epoch, stage, loss, acc_str, time))

    @staticmethod
    def update_name(hdf5, outdir):
        fname = os.path.join(outdir, hdf5)
        count = 0
        stem = hdf5.split('.')[0]
        while os.path.exists(fname) or os.path.exists(os.path.splitext(fname)[0] + '.npz'):
            count += 1
            fname = os.path.join(outdir, '{}_{:03d}.hdf5'.format(stem, count))
        return fname

    # ------------------------------------------------------------------ checkpoints (NeuralNet.py:768-825)
    def save_model(self, filename='model.pth.tar'):
        if self.engine is not None:
            self.model.load_state_dict(self.engine.state_dict())
        state = {'model': self.model.state_dict(), 'optimizer': self.optimizer.state_dict(),
                 'node': self.node_feature, 'edge': self.edge_feature, 'target': self.target, 'task': self.task,
                 'classes': self.classes, 'class_weight': self.class_weights, 'batch_size': self.batch_size,
                 'percent': self.percent, 'lr': self.lr, 'index': self.index, 'shuffle': self.shuffle,
                 'threshold': self.threshold, 'cluster_nodes': self.cluster_nodes,
                 'transform_sigmoid': self.transform_sigmoid}
        torch.save(state, filename)

    def load_params(self, filename):
        state = torch.load(filename, map_location='cpu', weights_only=False)
        self.node_feature, self.edge_feature, self.target = state['node'], state['edge'], state['target']
        self.batch_size, self.percent, self.lr, self.index = state['batch_size'], state['percent'], state['lr'], state['index']
        self.class_weights, self.task, self.classes = state['class_weight'], state['task'], state['classes']
        self.threshold, self.shuffle, self.cluster_nodes = state['threshold'], state['shuffle'], state['cluster_nodes']
        self.transform_sigmoid = state.get('transform_sigmoid', False)
        self.opt_loaded_state_dict = state['optimizer']
        self.model_load_state_dict = state['model']
