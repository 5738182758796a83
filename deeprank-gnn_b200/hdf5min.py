"""Minimal read-only HDF5 reader (no h5py in the target image).

Covers exactly the subset ``Graph.nx2h5`` (reference ``deeprank_gnn/Graph.py:61-139``)
and ``NeuralNet._export_epoch_hdf5`` (``NeuralNet.py:827-872``) produce with default
h5py settings: superblock v0, v1 object headers (with continuation blocks), groups as
symbol tables (v1 B-tree + SNOD + local heap), contiguous or compact dataset layout,
fixed-point / IEEE float / fixed-length string / enum datatypes, no chunking, no filters.

The API mimics the slice of h5py the host code needs::

    with File(path) as f:
        list(f.keys()); g = f['1ATN_1w']; 'score' in g
        a = g['node_data/bsa'][()]        # numpy array (or numpy scalar for rank 0)

When h5py is importable it is preferred by ``DataSet.open_hdf5``; this module is the
fallback, and the only path on the GPU box.
"""
import struct

import numpy as np

_SIG = b'\x89HDF\r\n\x1a\n'
_UNDEF = 0xFFFFFFFFFFFFFFFF


class HDF5FormatError(IOError):
    pass


class _Reader(object):
    def __init__(self, buf):
        self.buf = buf
        if buf[:8] != _SIG:
            raise HDF5FormatError('not an HDF5 file (bad signature)')
        ver = buf[8]
        if ver not in (0, 1):
            raise HDF5FormatError('unsupported superblock version %d (only v0/v1)' % ver)
        self.O = buf[13]
        self.L = buf[14]
        if self.O != 8 or self.L != 8:
            raise HDF5FormatError('only 8-byte offsets/lengths are supported')
        p = 24 if ver == 0 else 28
        self.base, _free, self.eof, _drv = struct.unpack_from('<4Q', buf, p)
        p += 32
        # root group symbol-table entry
        _name_off, self.root_header, cache_type = struct.unpack_from('<QQI', buf, p)
        self.root_scratch = buf[p + 24:p + 40]
        self.root_cache_type = cache_type

    # -- object headers -------------------------------------------------- #
    def messages(self, addr):
        """Yield (type, flags, payload bytes) of a version-1 object header."""
        buf = self.buf
        ver = buf[addr]
        if ver != 1:
            raise HDF5FormatError('unsupported object header version %d at 0x%x' % (ver, addr))
        nmsg, = struct.unpack_from('<H', buf, addr + 2)
        hsize, = struct.unpack_from('<I', buf, addr + 8)
        blocks = [(addr + 16, hsize)]
        out = []
        bi = 0
        while bi < len(blocks) and len(out) < nmsg:
            p, size = blocks[bi]
            end = p + size
            while p + 8 <= end and len(out) < nmsg:
                mtype, msize, mflags = struct.unpack_from('<HHB', buf, p)
                payload = buf[p + 8:p + 8 + msize]
                if mtype == 0x10:      # continuation
                    coff, clen = struct.unpack_from('<QQ', payload, 0)
                    blocks.append((coff, clen))
                out.append((mtype, mflags, payload))
                p += 8 + msize
            bi += 1
        return out

    # -- groups ---------------------------------------------------------- #
    def group_entries(self, header_addr):
        """dict name -> object header address for an old-style (symbol table) group."""
        stab = None
        for mtype, _f, payload in self.messages(header_addr):
            if mtype == 0x11:
                stab = struct.unpack_from('<QQ', payload, 0)
        if stab is None:
            return None
        btree, heap = stab
        buf = self.buf
        if buf[heap:heap + 4] != b'HEAP':
            raise HDF5FormatError('bad local heap signature')
        heap_data, = struct.unpack_from('<Q', buf, heap + 24)
        entries = {}

        def name_at(off):
            s = heap_data + off
            e = buf.index(b'\x00', s)
            return bytes(buf[s:e]).decode('utf8')

        def walk(node):
            if buf[node:node + 4] == b'TREE':
                _ntype, level, used = struct.unpack_from('<BBH', buf, node + 4)
                p = node + 24
                for i in range(used):
                    child, = struct.unpack_from('<Q', buf, p + 8 + i * 16)
                    walk(child)
            elif buf[node:node + 4] == b'SNOD':
                nsym, = struct.unpack_from('<H', buf, node + 6)
                p = node + 8
                for i in range(nsym):
                    noff, haddr = struct.unpack_from('<QQ', buf, p + i * 40)
                    entries[name_at(noff)] = haddr
            else:
                raise HDF5FormatError('unexpected group node signature at 0x%x' % node)

        walk(btree)
        return entries

    # -- datasets -------------------------------------------------------- #
    def _dtype(self, payload):
        cv = payload[0]
        cls, ver = cv & 0x0F, cv >> 4
        bits0 = payload[1]
        size, = struct.unpack_from('<I', payload, 4)
        order = '>' if (bits0 & 1) else '<'
        if cls == 0:
            signed = (bits0 >> 3) & 1
            return np.dtype('%s%s%d' % (order, 'i' if signed else 'u', size)), 8 + 4
        if cls == 1:
            return np.dtype('%sf%d' % (order, size)), 8 + 12
        if cls == 3:
            return np.dtype('S%d' % size), 8
        if cls == 8:
            base, _used = self._dtype(payload[8:])
            return base, None
        if cls == 9:
            return None, None
        raise HDF5FormatError('unsupported datatype class %d' % cls)

    def read_dataset(self, header_addr):
        shape = None
        dtype = None
        layout = None
        is_vlen = False
        for mtype, _f, payload in self.messages(header_addr):
            if mtype == 0x1:
                ver, rank, flags = payload[0], payload[1], payload[2]
                p = 8 if ver == 1 else 4
                shape = struct.unpack_from('<%dQ' % rank, payload, p) if rank else ()
            elif mtype == 0x3:
                dtype, _ = self._dtype(payload)
                is_vlen = dtype is None
            elif mtype == 0x8:
                ver = payload[0]
                if ver != 3:
                    raise HDF5FormatError('unsupported data layout version %d' % ver)
                cls = payload[1]
                if cls == 0:
                    size, = struct.unpack_from('<H', payload, 2)
                    layout = ('compact', payload[4:4 + size])
                elif cls == 1:
                    addr, size = struct.unpack_from('<QQ', payload, 2)
                    layout = ('contiguous', addr, size)
                else:
                    raise HDF5FormatError('chunked datasets are not supported by hdf5min')
        if shape is None or layout is None:
            raise HDF5FormatError('object at 0x%x is not a dataset' % header_addr)
        if is_vlen:
            raise HDF5FormatError('variable-length datatypes are not supported by hdf5min')
        count = int(np.prod(shape)) if len(shape) else 1
        if layout[0] == 'compact':
            raw = layout[1]
        else:
            _, addr, size = layout
            if addr == _UNDEF or count == 0:
                return np.zeros(shape, dtype=dtype.newbyteorder('='))
            raw = self.buf[addr:addr + count * dtype.itemsize]
        arr = np.frombuffer(raw, dtype=dtype, count=count).reshape(shape)
        arr = arr.astype(dtype.newbyteorder('='), copy=True)
        return arr[()] if arr.shape == () else arr


class _Node(object):
    def __init__(self, reader, addr, name):
        self._r, self._addr, self.name = reader, addr, name
        self._entries = None

    def _kids(self):
        if self._entries is None:
            self._entries = self._r.group_entries(self._addr)
        return self._entries


class Dataset(_Node):
    def __getitem__(self, key):
        arr = self._r.read_dataset(self._addr)
        if key == ():
            return arr
        return arr[key]

    @property
    def shape(self):
        return np.shape(self[()])


class Group(_Node):
    def keys(self):
        return sorted(self._kids().keys())     # h5py iterates symbol-table groups alphabetically

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self._kids())

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    def __getitem__(self, path):
        node = self
        for part in [p for p in path.split('/') if p]:
            kids = node._kids() if isinstance(node, Group) else None
            if kids is None or part not in kids:
                raise KeyError(path)
            addr = kids[part]
            child_entries = node._r.group_entries(addr)
            name = node.name.rstrip('/') + '/' + part
            if child_entries is None:
                node = Dataset(node._r, addr, name)
            else:
                node = Group(node._r, addr, name)
                node._entries = child_entries
        return node


class File(Group):
    def __init__(self, path, mode='r'):
        if mode != 'r':
            raise HDF5FormatError('hdf5min is read-only (install h5py to write HDF5)')
        with open(path, 'rb') as fh:
            buf = fh.read()
        reader = _Reader(memoryview(buf).toreadonly() if hasattr(memoryview, 'toreadonly') else buf)
        reader.buf = buf
        super().__init__(reader, reader.root_header, '/')
        self.filename = path

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
