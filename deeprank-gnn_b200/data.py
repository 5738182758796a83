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
            if k == '_structure':           # device-side cache of the structure pass: shared, read-only
                out.__dict__[k] = v
                continue
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
        self._c1_ptr = None        # int32 [B+1] segments of cluster1 (None without clusters)
        self._max_n = None         # host ints: largest graph (nodes / directed edges)
        self._max_e = None
        self._max_k0 = None        # host ints: most level-0 / level-1 clusters in one graph (None: unknown)
        self._max_k1 = None

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
        batch, node_ptr, edge_ptr, c1_ptr = [], [0], [0], [0]
        cum = 0
        has_c1 = 'cluster1' in keys
        max_k0 = max_k1 = 0
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
            if has_c1:
                c1_ptr.append(c1_ptr[-1] + d['cluster1'].numel())
                # exact per-graph cluster counts (shared-memory sizing of the per-graph fused kernels)
                max_k0 = max(max_k0, int(torch.unique(d['cluster0']).numel()))
                max_k1 = max(max_k1, int(torch.unique(d['cluster1']).numel()))
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
        out._c1_ptr = torch.tensor(c1_ptr, dtype=torch.int32) if has_c1 else None
        out._max_n = max(b - a for a, b in zip(node_ptr[:-1], node_ptr[1:]))
        out._max_e = max(b - a for a, b in zip(edge_ptr[:-1], edge_ptr[1:]))
        out._max_k0, out._max_k1 = (max_k0, max_k1) if has_c1 else (None, None)
        return out


def _pad4(n):
    return (n + 3) & ~3


class PackedBatch(object):
    """One mini-batch in a single contiguous host block (pinned when a GPU is present), ready
    for ONE host->device copy: the feeder-side replacement of ``data_batch.to(device)``
    (``NeuralNet.py:491``), which moves ~10 tensors separately.  Indices are int32.

    Sections (each padded to 16 bytes), all 4-byte elements viewed from one float32 buffer:
    ``x[N,F] edge_attr[E,ne] y[B] | edge_index[2,E] cluster0[N] node_ptr[B+1] edge_ptr[B+1]
    c1_ptr[B+1] | y_class[B] (int64, classification only) | cluster1[L1]``.  The only section
    whose length is not fixed by (B, N, E) comes last, so batches of equal (B, N, E) share
    every section offset (``layout_key``) and one captured CUDA graph serves them all.
    """
    FLOAT_SECTIONS = ('x', 'edge_attr', 'y')
    INT_SECTIONS = ('edge_index', 'cluster0', 'node_ptr', 'edge_ptr', 'c1_ptr')

    def __init__(self, B, N, E, L1, F, ne, max_n, max_e, with_class=False, max_k0=None, max_k1=None, idx16=False,
                 compact=False):
        self.B, self.N, self.E, self.L1, self.F, self.ne = B, N, E, L1, F, ne
        # idx16: edge_index travels as uint16 graph-LOCAL node ids (2E half-words = E words instead of 2E)
        self.idx16 = bool(idx16)
        # compact (needs idx16): only the FIRST half of every graph's directed edges travels ([2, E/2] uint16 - the
        # loader stores each edge in both directions, first half i -> j, second half j -> i, DataSet.py:266-269,
        # so the structure pass mirrors it back) and the cluster ids are uint16: E/2 + N/2 + L1/2 words instead of
        # E + N + L1 (cfg2: 1.96 -> 1.80 MB per step over PCIe)
        self.compact = bool(compact) and self.idx16
        if self.compact and E % 2:
            raise ValueError('compact records need an even number of directed edges')
        self.max_n, self.max_e = max_n, max_e
        # per-graph cluster-count bounds, rounded up so that batches of one shape share a layout key
        # (and therefore one captured CUDA graph) although their exact counts differ
        self.max_k0 = None if not max_k0 else min(max_n, (max_k0 + 31) // 32 * 32)
        self.max_k1 = None if not max_k1 else min(max_n, (max_k1 + 15) // 16 * 16)
        self.with_class = with_class
        half = lambda n: (n + 1) // 2
        sizes = dict(x=N * F, edge_attr=E * ne, y=B,
                     edge_index=(E // 2 if self.compact else E if self.idx16 else 2 * E),
                     cluster0=half(N) if self.compact else N, cluster1=half(L1) if self.compact else L1,
                     node_ptr=B + 1, edge_ptr=B + 1, c1_ptr=B + 1)
        self.offsets = {}
        o = 0
        for k in self.FLOAT_SECTIONS + self.INT_SECTIONS:
            self.offsets[k] = (o, sizes[k])
            o += _pad4(sizes[k])
        if with_class:
            self.offsets['y_class'] = (o, 2 * B)
            o += _pad4(2 * B)
        self.offsets['cluster1'] = (o, sizes['cluster1'])
        self.numel = o + _pad4(sizes['cluster1'])
        self.capacity_numel = o + _pad4(sizes['cluster0'])          # len(cluster1) <= N
        self.buf = None
        self.mol = None
        self.has_y = False

    def layout_key(self):
        return (self.B, self.N, self.E, self.F, self.ne, self.max_n, self.max_e, self.with_class, self.max_k0, self.max_k1,
                self.idx16, self.compact)

    @property
    def nbytes(self):
        return 4 * self.numel

    def views(self, buf, capacity=False):
        """Typed views of the sections inside ``buf`` (a float32 tensor of ``numel`` elements,
        host or device).  ``capacity=True`` (device staging buffers of ``capacity_numel`` elements
        only): ``cluster1`` is a view of N entries whose live length is ``L1``."""
        v = {}
        for k in self.FLOAT_SECTIONS:
            o, n = self.offsets[k]
            v[k] = buf[o:o + n]
        ibuf = buf.view(torch.int32)
        for k in self.INT_SECTIONS + ('cluster1',):
            o, n = self.offsets[k]
            v[k] = ibuf[o:o + n]
        if capacity:                                # staging buffers: capacity-sized view, live length = L1
            if buf.numel() < self.capacity_numel:
                raise ValueError('a capacity view needs a buffer of capacity_numel = %d elements' % self.capacity_numel)
            o, _n = self.offsets['cluster1']
            v['cluster1'] = ibuf[o:o + self.N]
        v['x'] = v['x'].view(self.N, self.F)
        v['edge_attr'] = v['edge_attr'].view(self.E, self.ne) if self.ne else None
        if self.compact:                            # uint16 sections (torch has no uint16: int16 views, ids <= 32767)
            hbuf = buf.view(torch.int16)
            o, _n = self.offsets['edge_index']
            v['edge_index'] = hbuf[2 * o:2 * o + self.E].view(2, self.E // 2)
            o, _n = self.offsets['cluster0']
            v['cluster0'] = hbuf[2 * o:2 * o + self.N]
            o, _n = self.offsets['cluster1']
            v['cluster1'] = hbuf[2 * o:2 * o + (self.N if capacity else self.L1)]
        elif self.idx16:
            o, _n = self.offsets['edge_index']
            v['edge_index'] = buf.view(torch.int16)[2 * o:2 * o + 2 * self.E].view(2, self.E)
        else:
            v['edge_index'] = v['edge_index'].view(2, self.E)
        if self.with_class:
            o, n = self.offsets['y_class']
            v['y_class'] = ibuf[o:o + n].view(torch.int64)
        else:
            v['y_class'] = None
        return v

    @staticmethod
    def _mirrored_halves(batch):
        """Boolean mask of the edges in the FIRST half of their graph's edge list when every graph stores its
        edges as [i -> j pairs | the same pairs j -> i] (the loader's layout, DataSet.py:266-269); None otherwise."""
        ep = batch._edge_ptr.long()
        counts = ep[1:] - ep[:-1]
        if int((counts % 2).sum()) != 0:
            return None
        E = batch.edge_index.size(1)
        if E == 0:
            return torch.zeros(0, dtype=torch.bool)
        start = torch.repeat_interleave(ep[:-1], counts)
        half = torch.repeat_interleave(counts // 2, counts)
        pos = torch.arange(E) - start
        first = pos < half
        ei = batch.edge_index
        a, b = ei[:, first], ei[:, ~first]         # per graph, both keep their order: pair k of a graph <-> pair k
        if not (torch.equal(a[0], b[1]) and torch.equal(a[1], b[0])):
            return None
        return first

    @staticmethod
    def from_batch(batch, pin=None, classes=None, edge_attr=True, idx16=False, compact=None):
        """Pack a collated ``Batch``.  ``classes``: for classification, the class list used to
        map targets to class indices (``format_output``, NeuralNet.py:616-631).  Compact options that
        cut the bytes a step moves over PCIe: ``edge_attr=False`` leaves the edge attributes out
        (GINet's attention is the identity - alpha == 1, SURVEY a1 - and FoutNet ignores them; only sGAT
        reads them), ``idx16=True`` stores ``edge_index`` as uint16 graph-local node ids; ``compact`` (default: with
        ``idx16``, whenever the batch allows it) additionally sends only the first half of every graph's mirrored
        edge list and the cluster ids as uint16."""
        if batch._node_ptr is None or batch._c1_ptr is None:
            raise ValueError('PackedBatch needs a Batch collated by Batch.from_data_list with cluster0/cluster1')
        x = batch.x
        ea = getattr(batch, 'edge_attr', None)
        if ea is not None and ea.dim() == 1:
            ea = ea.unsqueeze(-1)
        N, F = x.size(0), (x.size(1) if x.dim() == 2 else 1)
        E = batch.edge_index.size(1)
        if not edge_attr:
            ea = None
        ne = 0 if ea is None else ea.size(1)
        idx16 = bool(idx16) and batch._max_n <= 32767
        half_mask = None
        if idx16 and (compact is None or compact):
            c0, c1 = batch.cluster0, batch.cluster1
            ids_ok = (c0.numel() == 0 or (int(c0.min()) >= 0 and int(c0.max()) <= 32767)) and \
                (c1.numel() == 0 or (int(c1.min()) >= 0 and int(c1.max()) <= 32767))
            half_mask = PackedBatch._mirrored_halves(batch) if ids_ok else None
        pb = PackedBatch(batch.num_graphs, N, E, batch.cluster1.numel(), F, ne, batch._max_n, batch._max_e,
                         with_class=classes is not None, max_k0=batch._max_k0, max_k1=batch._max_k1, idx16=idx16,
                         compact=half_mask is not None)
        pin = torch.cuda.is_available() if pin is None else pin
        pb.buf = torch.zeros(pb.numel, dtype=torch.float32, pin_memory=bool(pin))
        v = pb.views(pb.buf)
        v['x'].copy_(x.reshape(N, F))
        if ne:
            v['edge_attr'].copy_(ea)
        y = getattr(batch, 'y', None)
        if y is not None:
            pb.has_y = True
            v['y'].copy_(y.reshape(-1))
            if classes is not None:
                c2i = {int(c): i for i, c in enumerate(classes)}
                v['y_class'].copy_(torch.tensor([c2i[int(t)] for t in y.reshape(-1).tolist()], dtype=torch.int64))
        if idx16:
            counts = (batch._edge_ptr[1:] - batch._edge_ptr[:-1]).long()
            first = torch.repeat_interleave(batch._node_ptr[:-1].long(), counts)     # first node of every edge's graph
            local = batch.edge_index - first.unsqueeze(0)
            v['edge_index'].copy_(local[:, half_mask] if pb.compact else local)
        else:
            v['edge_index'].copy_(batch.edge_index)
        v['cluster0'].copy_(batch.cluster0)
        v['cluster1'].copy_(batch.cluster1)
        v['node_ptr'].copy_(batch._node_ptr)
        v['edge_ptr'].copy_(batch._edge_ptr)
        v['c1_ptr'].copy_(batch._c1_ptr)
        pb.mol = getattr(batch, 'mol', None)
        return pb


class PackedCache(object):
    """Packed, memory-mapped on-disk cache of feeder records: every mini-batch of a pass is
    stored as its ``PackedBatch`` block (compact: uint16 graph-local edge ids, int32 cluster ids, fp32 features)
    in ONE binary file, 64-byte aligned, behind a small JSON index.  An epoch then never touches HDF5 or Python
    collation (``DataSet.py:241-366`` opens the file and rebuilds ~10 tensors per graph): ``cache[i]`` is a
    ``PackedBatch`` whose buffer is a view of the memory map (``pin=True``: a pinned copy, ready for the
    asynchronous host->device copy of ``Engine.train_batches``).

    File: ``b'DRGNNPC1'`` | uint64 index length | JSON index | padding to 64 | records."""

    MAGIC = b'DRGNNPC1'
    FIELDS = ('B', 'N', 'E', 'L1', 'F', 'ne', 'max_n', 'max_e')

    @staticmethod
    def build(path, packed_batches):
        """Write ``packed_batches`` (an iterable of ``PackedBatch``) to ``path``.  Returns the number of records."""
        import json
        metas, blobs, off = [], [], 0
        for pb in packed_batches:
            m = {k: int(getattr(pb, k)) for k in PackedCache.FIELDS}
            m.update(with_class=bool(pb.with_class), idx16=bool(pb.idx16), compact=bool(pb.compact), has_y=bool(pb.has_y),
                     max_k0=pb.max_k0, max_k1=pb.max_k1, numel=int(pb.numel), offset=off,
                     mol=list(pb.mol) if pb.mol is not None else None)
            metas.append(m)
            blobs.append(pb.buf.detach().cpu().numpy().view(np.uint8))
            off += (4 * pb.numel + 63) // 64 * 64
        index = json.dumps({'version': 1, 'records': metas}).encode()
        head = PackedCache.MAGIC + np.uint64(len(index)).tobytes() + index
        head += b'\0' * ((-len(head)) % 64)
        with open(path, 'wb') as f:
            f.write(head)
            for m, raw in zip(metas, blobs):
                f.write(raw.tobytes())
                f.write(b'\0' * ((-raw.nbytes) % 64))
        return len(metas)

    def __init__(self, path, pin=False, register=False):
        """``pin=True``: every record is copied into its own pinned block when read.  ``register=True``: the
        memory map itself is page-locked once (``cudaHostRegister``), so records are handed to the asynchronous
        host->device copy of ``Engine.train_batches`` / ``drgnn_feed_run`` straight from the mapping - no copy, no
        per-record pinned allocation (falls back to ``pin=True`` when the driver refuses the mapping)."""
        import json
        self.path, self.pin = path, bool(pin)
        self.registered = False
        with open(path, 'rb') as f:
            if f.read(8) != self.MAGIC:
                raise ValueError('%s is not a packed cache' % path)
            n = int(np.frombuffer(f.read(8), dtype=np.uint64)[0])
            index = json.loads(f.read(n).decode())
        if index.get('version') != 1:
            raise ValueError('unsupported packed-cache version %r' % index.get('version'))
        self.records = index['records']
        self._base = (16 + n + 63) // 64 * 64
        # 'c' (private copy-on-write): page-lockable without the read-only registration flag; nothing is written
        self._map = np.memmap(path, dtype=np.float32, mode='c' if register else 'r', offset=self._base) \
            if self.records else None
        self._tmap = None
        if register and self._map is not None:
            self._register()

    def _register(self):
        try:
            t = torch.from_numpy(self._map)
            rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * 4, 0)
            if int(rc) != 0:
                raise RuntimeError('cudaHostRegister returned %s' % rc)
            self._tmap, self.registered = t, True
        except Exception:
            self.pin = True               # per-record pinned copies instead

    def close(self):
        if self.registered and self._tmap is not None:
            try:
                torch.cuda.cudart().cudaHostUnregister(self._tmap.data_ptr())
            except Exception:
                pass
            self.registered, self._tmap = False, None

    def __del__(self):
        self.close()

    def __len__(self):
        return len(self.records)

    def __getitem__(self, i):
        m = self.records[i]
        pb = PackedBatch(m['B'], m['N'], m['E'], m['L1'], m['F'], m['ne'], m['max_n'], m['max_e'],
                         with_class=m['with_class'], idx16=m['idx16'], compact=m.get('compact', False))
        pb.max_k0, pb.max_k1 = m['max_k0'], m['max_k1']          # stored already rounded
        assert pb.numel == m['numel'], 'packed-cache record %d does not match the record layout' % i
        start = m['offset'] // 4
        if self.registered:
            view = self._tmap[start:start + pb.numel]            # page-locked mapping: DMA source as it is
        else:
            view = torch.from_numpy(np.asarray(self._map[start:start + pb.numel]))
        if self.pin and not self.registered:
            buf = torch.empty(pb.numel, dtype=torch.float32, pin_memory=True)
            buf.copy_(view)
        else:
            buf = view                                            # read-only view of the memory map
        pb.buf, pb.has_y, pb.mol = buf, m['has_y'], m['mol']
        return pb

    def __iter__(self):
        return (self[i] for i in range(len(self)))


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
