"""Graph records and mini-batch collation (host side).

The reference builds ``torch_geometric.data.Data`` records in
``HDF5DataSet.load_one_graph`` (``deeprank_gnn/DataSet.py:231-366``) and collates them
with PyG's ``DataLoader`` / ``Batch.from_data_list`` (``NeuralNet.py:153-154``).
torch_geometric is not a dependency here; these classes keep the attribute names and
the collation rule (keys containing ``index`` are concatenated on the last dim and
offset by the cumulative node count, everything else on dim 0, ``batch`` = graph id
per node) and additionally record the CSR-of-graphs pointers the CUDA structure
kernels need (``node_ptr``, ``edge_ptr``), so no device-side scan / host sync is
needed to find graph boundaries.
"""
import copy

import numpy as np
import torch


class Data(object):
    """One graph.  Same attribute names as the reference record."""

    def __init__(self, x=None, edge_index=None, edge_attr=None, y=None, pos=None, **kwargs):
        self.x = x
        self.edge_index = edge_index
        self.edge_attr = edge_attr
        self.y = y
        self.pos = pos
        for k, v in kwargs.items():
            setattr(self, k, v)

    # -- PyG-like introspection ---------------------------------------- #
    @property
    def keys(self):
        return [k for k, v in self.__dict__.items() if v is not None and not k.startswith('_')]

    def __contains__(self, key):
        return key in self.keys

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    @property
    def num_nodes(self):
        if self.x is not None:
            return self.x.size(0)
        if self.pos is not None:
            return self.pos.size(0)
        return int(self.edge_index.max()) + 1

    @property
    def num_edges(self):
        return 0 if self.edge_index is None else self.edge_index.size(1)

    @property
    def num_features(self):
        if self.x is None:
            return 0
        return 1 if self.x.dim() == 1 else self.x.size(1)

    num_node_features = num_features

    def clone(self):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = v.clone() if torch.is_tensor(v) else copy.deepcopy(v)
        return out

    def to(self, device, non_blocking=False):
        for k, v in self.__dict__.items():
            if torch.is_tensor(v):
                self.__dict__[k] = v.to(device, non_blocking=non_blocking)
        return self

    def pin_memory(self):
        for k, v in self.__dict__.items():
            if torch.is_tensor(v) and v.device.type == 'cpu':
                self.__dict__[k] = v.pin_memory()
        return self

    def __repr__(self):
        parts = []
        for k in self.keys:
            v = self[k]
            parts.append('%s=%s' % (k, list(v.shape) if torch.is_tensor(v) else type(v).__name__))
        return '%s(%s)' % (self.__class__.__name__, ', '.join(parts))


class Batch(Data):
    """A block-diagonal mini-batch of independent graphs."""

    def __init__(self, batch=None, **kwargs):
        super().__init__(**kwargs)
        self.batch = batch
        self._num_graphs = None
        self._node_ptr = None      # int32 [B+1], same device as the tensors
        self._edge_ptr = None      # int32 [B+1]

    @property
    def num_graphs(self):
        if self._num_graphs is None:
            if self.batch is None:
                return 1
            self._num_graphs = int(self.batch.max()) + 1      # host sync (hand-made batches only)
        return self._num_graphs

    @staticmethod
    def from_data_list(data_list):
        if len(data_list) == 0:
            raise ValueError('cannot collate an empty list of graphs')
        keys = data_list[0].keys
        out = Batch()
        cols = {k: [] for k in keys}
        batch, node_ptr, edge_ptr = [], [0], [0]
        cum = 0
        for i, d in enumerate(data_list):
            n = d.num_nodes
            for k in keys:
                v = d[k]
                if torch.is_tensor(v) and 'index' in k:
                    v = v + cum
                cols[k].append(v)
            batch.append(torch.full((n,), i, dtype=torch.long))
            cum += n
            node_ptr.append(cum)
            edge_ptr.append(edge_ptr[-1] + d.num_edges)
        for k in keys:
            items = cols[k]
            if torch.is_tensor(items[0]):
                out.__dict__[k] = torch.cat(items, dim=-1 if 'index' in k else 0)
            else:
                out.__dict__[k] = items
        out.batch = torch.cat(batch, dim=0)
        out._num_graphs = len(data_list)
        out._node_ptr = torch.tensor(node_ptr, dtype=torch.int32)
        out._edge_ptr = torch.tensor(edge_ptr, dtype=torch.int32)
        return out


class DataLoader(object):
    """Worker-less loader with the call signature the reference uses
    (``DataLoader(dataset, batch_size=..., shuffle=...)``, ``NeuralNet.py:105,153,158``).
    ``dataset`` needs ``__len__``/``len()`` and ``get(i)`` (or ``__getitem__``)."""

    def __init__(self, dataset, batch_size=1, shuffle=False, pin_memory=False, generator=None):
        self.dataset = dataset
        self.batch_size = batch_size
        self.shuffle = shuffle
        self.pin_memory = pin_memory
        self.generator = generator

    def _len_dataset(self):
        return self.dataset.len() if hasattr(self.dataset, 'len') else len(self.dataset)

    def __len__(self):
        n = self._len_dataset()
        return (n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = self._len_dataset()
        if self.shuffle:
            order = torch.randperm(n, generator=self.generator).tolist()
        else:
            order = list(range(n))
        get = self.dataset.get if hasattr(self.dataset, 'get') else self.dataset.__getitem__
        for s in range(0, n, self.batch_size):
            graphs = [get(i) for i in order[s:s + self.batch_size]]
            graphs = [g for g in graphs if g is not None]
            if not graphs:
                continue
            b = Batch.from_data_list(graphs)
            if self.pin_memory:
                b.pin_memory()
            yield b
