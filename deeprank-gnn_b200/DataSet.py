"""Dataset side of the drop-in boundary: ``HDF5DataSet``, ``DivideDataSet``, ``PreCluster``.

Same constructor arguments, attributes (``index_complexes``, ``node_feature``,
``edge_feature``, ``database``) and record layout as the reference
(``deeprank_gnn/DataSet.py:14-42, 45-88, 91-450``; record = SURVEY 8a row a14), but

* files are parsed with h5py when it is importable, else with the bundled read-only
  ``hdf5min`` reader (h5py is not part of the target image);
* a file is opened once and kept open (the reference opens and closes the HDF5 file for
  every graph access, ``DataSet.py:241,365``) and decoded graphs are memoised, so epochs
  after the first never touch the file again;
* ``PreCluster`` needs community detection (``markov_clustering`` / ``python-louvain``),
  an offline preprocessing step outside the hot path: clusters already stored in the file
  are used as they are; when they are missing and no detection backend is importable a
  clear error is raised instead of silently training without pooling.
"""
import copy
import operator
import re
import sys

import numpy as np
import torch

from .data import Data

try:                                    # pragma: no cover - not available in the target image
    import h5py as _h5py
except Exception:                       # noqa
    _h5py = None
from . import hdf5min


def open_hdf5(path):
    """Read-only handle with the slice of the h5py API the loaders use."""
    if _h5py is not None:
        return _h5py.File(path, 'r')
    return hdf5min.File(path)


def _default_edge_transform(x):
    return np.tanh(-x / 2 + 2) + 1      # DataSet.py:96


_COND = re.compile(r'\s*(<=|>=|==|!=|<|>)\s*([-+0-9.eE]+)\s*')
_OPS = {'<': operator.lt, '>': operator.gt, '==': operator.eq, '<=': operator.le, '>=': operator.ge,
        '!=': operator.ne}


def DivideDataSet(dataset, percent=[0.8, 0.2], shuffle=True):
    """Split into a training and an evaluation set (DataSet.py:14-42): numpy shuffle of the
    indices, first ``int(percent[0]*size)`` graphs train, the rest eval."""
    size = dataset.len()
    index = np.arange(size)
    if shuffle:
        np.random.shuffle(index)
    cut = int(percent[0] * size)
    parts = []
    for sel in (index[:cut], index[cut:]):
        part = copy.copy(dataset)                 # shares the open files and the graph cache
        part.index_complexes = [dataset.index_complexes[i] for i in sel]
        parts.append(part)
    return parts[0], parts[1]


def PreCluster(dataset, method):
    """Make sure every graph carries ``cluster0`` / ``cluster1`` for ``method`` (DataSet.py:45-88).

    The reference recomputes MCL / Louvain clusters for every graph on every ``NeuralNet``
    construction and rewrites the HDF5 file in place.  Community detection is offline
    preprocessing: here clusters already stored under
    ``clustering/<method>/depth_{0,1}`` are authoritative; graphs without them are clustered
    with ``community_pooling.community_detection`` (needs ``markov_clustering`` or
    ``python-louvain``) and the result is kept in memory (the file is not rewritten)."""
    method = method.lower()
    missing = []
    for i, (fname, mol) in enumerate(dataset.index_complexes):
        data = dataset.load_one_graph(fname, mol)
        if data is None:
            continue
        if getattr(data, 'cluster0', None) is None or getattr(data, 'cluster1', None) is None:
            missing.append((fname, mol))
    if not missing:
        return
    from .community_pooling import community_detection, community_pooling_host, mcl_detection_batch
    graphs = [dataset.load_one_graph(fname, mol) for fname, mol in missing]
    if method == 'mcl' and torch.cuda.is_available():
        # GPU pre-clustering (csrc/mcl.cu): both levels of ALL missing graphs in two launches, one CTA per graph
        c0s = mcl_detection_batch([g.internal_edge_index for g in graphs], [g.num_nodes for g in graphs])
        pooled = [community_pooling_host(c0, g) for c0, g in zip(c0s, graphs)]
        c1s = mcl_detection_batch([p.internal_edge_index for p in pooled], [p.num_nodes for p in pooled])
    else:
        c0s, c1s = [], []
        for data in graphs:
            c0 = community_detection(data.internal_edge_index, data.num_nodes, method=method)
            p = community_pooling_host(c0, data)
            c0s.append(c0)
            c1s.append(community_detection(p.internal_edge_index, p.num_nodes, method=method))
    for (fname, mol), data, c0, c1 in zip(missing, graphs, c0s, c1s):
        data.cluster0, data.cluster1 = c0, c1
        dataset._cache[(fname, mol)] = data


class HDF5DataSet(object):
    def __init__(self, root='./', database=None, transform=None, pre_transform=None, dict_filter=None, target=None,
                 tqdm=True, index=None, node_feature='all', edge_feature=['dist'], clustering_method='mcl',
                 edge_feature_transform=_default_edge_transform, verbose=False):
        self.root = root
        self.transform = transform
        self.pre_transform = pre_transform
        self.database = list(database) if isinstance(database, (list, tuple)) else [database]
        self.target = target
        self.dict_filter = dict_filter
        self.tqdm = tqdm
        self.index = index
        self.node_feature = node_feature
        self.edge_feature = edge_feature
        self.edge_feature_transform = edge_feature_transform
        self.clustering_method = clustering_method
        self.verbose = verbose
        self._files = {}
        self._cache = {}
        self.check_hdf5_files()
        self.check_node_feature()
        self.check_edge_feature()
        self.create_index_molecules()

    # ------------------------------------------------------------------ plumbing
    def _log(self, *a):
        if self.verbose:
            print(*a)
            sys.stdout.flush()

    def _file(self, fname):
        f = self._files.get(fname)
        if f is None:
            f = open_hdf5(fname)
            self._files[fname] = f
        return f

    def __copy__(self):
        new = self.__class__.__new__(self.__class__)
        new.__dict__.update(self.__dict__)
        return new

    def __deepcopy__(self, memo):
        new = self.__copy__()
        new.index_complexes = list(self.index_complexes)
        return new

    def len(self):
        return len(self.index_complexes)

    __len__ = len

    def get(self, index):
        fname, mol = self.index_complexes[index]
        data = self.load_one_graph(fname, mol)
        if data is not None and self.transform is not None:
            data = self.transform(data)
        return data

    __getitem__ = get

    # ------------------------------------------------------------------ checks (DataSet.py:169-229)
    def check_hdf5_files(self):
        self._log('   Checking dataset Integrity')
        bad = []
        for fname in self.database:
            try:
                if len(list(self._file(fname).keys())) == 0:
                    self._log('    -> %s is empty ' % fname)
                    bad.append(fname)
            except Exception as e:      # corrupted / unreadable files are dropped like the reference does
                self._log(e)
                self._log('    -> %s is corrupted ' % fname)
                bad.append(fname)
        for fname in bad:
            self.database.remove(fname)
            self._files.pop(fname, None)
        if not self.database:
            raise ValueError('no readable HDF5 file in the database')

    def _first_group(self):
        f = self._file(self.database[0])
        return f[list(f.keys())[0]]

    def check_node_feature(self):
        self.available_node_feature = list(self._first_group()['node_data'].keys())
        if self.node_feature == 'all':
            self.node_feature = self.available_node_feature
        else:
            for feat in self.node_feature:
                if feat not in self.available_node_feature:
                    raise ValueError('node feature %r not found in %s; available: %s'
                                     % (feat, self.database[0], ', '.join(self.available_node_feature)))

    def check_edge_feature(self):
        self.available_edge_feature = list(self._first_group()['edge_data'].keys())
        if self.edge_feature == 'all':
            self.edge_feature = self.available_edge_feature
        elif self.edge_feature is not None:
            for feat in self.edge_feature:
                if feat not in self.available_edge_feature:
                    raise ValueError('edge feature %r not found in %s; available: %s'
                                     % (feat, self.database[0], ', '.join(self.available_edge_feature)))

    # ------------------------------------------------------------------ one record (DataSet.py:231-366)
    def _stack(self, grp, prefix, feats):
        cols = []
        for feat in feats:
            vals = np.asarray(grp[prefix + feat][()])
            cols.append(vals.reshape(-1, 1) if vals.ndim == 1 else vals)
        return np.hstack(cols)

    def _both_directions(self, ind):
        ind = np.asarray(ind).reshape(-1, 2)
        return torch.tensor(np.ascontiguousarray(np.vstack((ind, np.flip(ind, 1))).T), dtype=torch.long)

    def _edge_attr(self, grp, prefix):
        if self.edge_feature is None:
            return None
        vals = self._stack(grp, prefix, self.edge_feature)
        vals = self.edge_feature_transform(np.vstack((vals, vals)))
        return torch.tensor(vals, dtype=torch.float).contiguous()

    def load_one_graph(self, fname, mol):
        key = (fname, mol)
        if key in self._cache:
            return self._cache[key].clone()
        try:
            grp = self._file(fname)[mol]
        except Exception:
            return None
        try:
            x = torch.tensor(self._stack(grp, 'node_data/', self.node_feature), dtype=torch.float)
        except Exception:
            self._log('node attributes not found in the file', fname)
            return None
        try:
            edge_index = self._both_directions(grp['edge_index'][()])
            edge_attr = self._edge_attr(grp, 'edge_data/')
            internal_edge_index = self._both_directions(grp['internal_edge_index'][()])
            internal_edge_attr = self._edge_attr(grp, 'internal_edge_data/')
        except Exception:
            self._log('edge features not found in the file', fname)
            return None
        y = None
        if self.target is not None and self.target in grp['score'].keys():
            val = grp['score/' + self.target][()]
            if val is not None:
                y = torch.tensor([float(val)], dtype=torch.float)
        pos = torch.tensor(np.asarray(grp['node_data/pos'][()]), dtype=torch.float).contiguous()
        data = Data(x=x, edge_index=edge_index, edge_attr=edge_attr, y=y, pos=pos)
        data.internal_edge_index = internal_edge_index
        data.internal_edge_attr = internal_edge_attr
        data.mol = mol
        cpath = 'clustering/%s' % self.clustering_method
        found = False
        if 'clustering' in grp.keys() and self.clustering_method in grp['clustering'].keys():
            cg = grp[cpath]
            if 'depth_0' in cg.keys() and 'depth_1' in cg.keys():
                data.cluster0 = torch.tensor(np.asarray(cg['depth_0'][()]), dtype=torch.long)
                data.cluster1 = torch.tensor(np.asarray(cg['depth_1'][()]), dtype=torch.long)
                found = True
        if not found:
            self._log('WARNING: no cluster detected')
        self._cache[key] = data
        return data.clone()

    # ------------------------------------------------------------------ index (DataSet.py:368-450)
    def create_index_molecules(self):
        self._log('   Processing data set')
        self.index_complexes = []
        for fdata in self.database:
            try:
                f = self._file(fdata)
                names = list(f.keys())
                if self.index is not None:
                    names = [names[i] for i in self.index]
                for k in names:
                    if self.filter(f[k]):
                        self.index_complexes.append((fdata, k))
            except Exception as inst:
                self._log('\t\t--> Ignore File : ' + fdata)
                self._log(inst)
        self.ntrain = len(self.index_complexes)
        self.index_train = list(range(self.ntrain))
        self.ntot = len(self.index_complexes)

    def filter(self, molgrp):
        """``dict_filter = {'irmsd': '<10', ...}``; several conditions on one score may be
        joined with ``and`` / ``or`` (the reference evals the string, DataSet.py:409-450)."""
        if self.dict_filter is None:
            return True
        for name, cond in self.dict_filter.items():
            try:
                val = float(molgrp['score'][name][()])
            except KeyError:
                raise ValueError('filter %r not found; options: %s' % (name, ', '.join(molgrp['score'].keys())))
            if not isinstance(cond, str):
                raise ValueError('Conditions not supported', cond)
            if not _eval_condition(cond, val):
                return False
        return True


def _eval_condition(cond, val):
    ors = []
    for disj in re.split(r'\bor\b', cond):
        ands = []
        for term in re.split(r'\band\b', disj):
            m = _COND.fullmatch(term)
            if m is None:
                raise ValueError('Conditions not supported', cond)
            ands.append(_OPS[m.group(1)](val, float(m.group(2))))
        ors.append(all(ands))
    return any(ors)
