"""CPU tests of the host side: C-ABI exports, record / batch / packed-batch logic, HDF5 reader,
dataset API, parameter layout, sharding and the 2-rank (gloo) gradient all-reduce contract."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import FIXTURE, ROOT


def test_library_exports_every_declared_symbol(lib):
    """libdrgnn.so loads and exports exactly what include/drgnn.h declares (no compute call)."""
    hdr = open(os.path.join(ROOT, 'include', 'drgnn.h')).read()
    declared = set(re.findall(r'\b(drgnn_[a-z0-9_]+)\s*\(', hdr))
    assert len(declared) >= 20
    from deeprank_gnn_b200 import _lib
    assert declared == set(_lib.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.drgnn_version() == 100
    out = subprocess.run(['nm', '-D', '--defined-only', _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r'\bT (drgnn_[a-z0-9_]+)', out))
    assert declared <= exported


def test_ctypes_structs_match_header_layout(lib):
    """Field counts / order of the ctypes mirrors follow the header (a mismatch would shift pointers)."""
    from deeprank_gnn_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'drgnn.h')).read()

    def fields(struct):
        body = hdr[hdr.index('typedef struct ' + struct):]
        body = body[body.index('{') + 1:body.index('} ' + struct)]
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        names = []
        for decl in body.split(';'):
            decl = decl.strip()
            decl = re.sub(r'\[[^\]]*\]\s*$', '', decl)          # array members: name[DRGNN_MAX_PEERS]
            if decl:
                names.append(re.findall(r'([A-Za-z0-9_]+)\s*$', decl)[0])
        return names

    assert fields('drgnn_structure_io') == [f[0] for f in _lib.StructureIO._fields_]
    assert fields('drgnn_aggregate_args') == [f[0] for f in _lib.AggregateArgs._fields_]
    assert fields('drgnn_linear_args') == [f[0] for f in _lib.LinearArgs._fields_]
    assert fields('drgnn_linear_wgrad_args') == [f[0] for f in _lib.LinearWgradArgs._fields_]
    assert fields('drgnn_head_args') == [f[0] for f in _lib.HeadArgs._fields_]
    assert fields('drgnn_ginet_fused_args') == [f[0] for f in _lib.GinetFusedArgs._fields_]
    assert fields('drgnn_ginet_step_args') == [f[0] for f in _lib.GinetStepArgs._fields_]
    assert fields('drgnn_peer_comm') == [f[0] for f in _lib.PeerComm._fields_]
    assert fields('drgnn_peer_adam_args') == [f[0] for f in _lib.PeerAdamArgs._fields_]
    assert fields('drgnn_feed_step') == [f[0] for f in _lib.FeedStep._fields_]
    import ctypes
    assert ctypes.sizeof(_lib.FeedStep) == 72          # Engine._feed_native fills the records as 9 x int64


def test_product_refuses_cpu_tensors(lib):
    from deeprank_gnn_b200 import ops
    from deeprank_gnn_b200._lib import DrgnnError
    x = torch.zeros(4, 4)
    with pytest.raises(DrgnnError):
        ops.aggregate(x, torch.zeros(5, dtype=torch.int32), torch.zeros(0, dtype=torch.int32), torch.zeros(4, 4))
    from deeprank_gnn_b200.engine import Engine
    with pytest.raises(DrgnnError):
        Engine('GINet', 8, device='cpu')


def test_hdf5_reader_reads_fixture_values():
    from deeprank_gnn_b200 import hdf5min
    with hdf5min.File(FIXTURE) as f:
        names = list(f.keys())
        assert len(names) == 10 and names[0] == '1ATN_10w'
        g = f['1ATN_1w']
        assert sorted(g['node_data'].keys()) == ['bsa', 'chain', 'charge', 'cons', 'depth', 'hse', 'ic', 'polarity',
                                                 'pos', 'pssm', 'type']
        assert g['node_data/pssm'][()].shape == (132, 20) and g['node_data/hse'][()].shape == (132, 3)
        assert g['edge_index'][()].shape == (374, 2) and g['edge_index'][()].dtype == np.int64
        assert abs(float(g['score/irmsd'][()]) - 14.919) < 1e-9
        assert g['clustering/mcl/depth_0'][()].shape == (132,)
        d = g['edge_data/dist'][()]
        assert d.dtype == np.float64 and 0 < d.min() and d.max() <= 8.5 + 1e-6


def test_dataset_api_and_filter():
    from deeprank_gnn_b200.DataSet import DivideDataSet, HDF5DataSet, PreCluster
    ds = HDF5DataSet(database=FIXTURE, node_feature=['type', 'polarity', 'bsa', 'depth', 'hse', 'ic', 'pssm'],
                     target='irmsd')
    assert ds.len() == 10 and ds.get(0).num_features == 28            # tests/test_nn.py feature list -> F = 28
    PreCluster(ds, 'mcl')                                             # clusters are stored in the fixture
    tr, va = DivideDataSet(ds, percent=[0.8, 0.2])
    assert tr.len() == 8 and va.len() == 2
    assert HDF5DataSet(database=FIXTURE, dict_filter={'irmsd': '<10'}, target='irmsd').len() == 0
    assert HDF5DataSet(database=FIXTURE, dict_filter={'irmsd': '>15 and <16'}, target='irmsd').len() == 6
    sub = HDF5DataSet(database=FIXTURE, index=[0, 2], target='fnat')
    assert [m for _, m in sub.index_complexes] == ['1ATN_10w', '1ATN_2w']
    with pytest.raises(ValueError):
        HDF5DataSet(database=FIXTURE, node_feature=['nope'])


def test_batch_collation_and_packed_roundtrip():
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.data import Batch, DataLoader, PackedBatch
    graphs = synthetic.make_graphs('cfg2', count=5, seed=3)
    b = Batch.from_data_list(graphs)
    assert b.num_graphs == 5 and b._node_ptr.tolist() == [0, 200, 400, 600, 800, 1000]
    assert b._edge_ptr.tolist() == [0, 1000, 2000, 3000, 4000, 5000]
    assert b._c1_ptr[-1] == b.cluster1.numel() and b._max_n == 200 and b._max_e == 1000
    assert torch.equal(b.edge_index[:, 1000:2000], graphs[1].edge_index + 200)
    assert torch.equal(b.cluster0[200:400], graphs[1].cluster0)                 # clusters are not offset
    pb = PackedBatch.from_batch(b, pin=False)
    v = pb.views(pb.buf)
    assert torch.equal(v['x'], b.x) and torch.equal(v['edge_attr'], b.edge_attr) and torch.equal(v['y'], b.y)
    assert torch.equal(v['edge_index'].long(), b.edge_index) and torch.equal(v['cluster1'].long(), b.cluster1)
    assert torch.equal(v['c1_ptr'], b._c1_ptr)
    other = PackedBatch.from_batch(Batch.from_data_list(synthetic.make_graphs('cfg2', count=5, seed=9)), pin=False)
    assert other.layout_key() == pb.layout_key() and other.capacity_numel == pb.capacity_numel
    for k in PackedBatch.FLOAT_SECTIONS + PackedBatch.INT_SECTIONS:
        assert other.offsets[k][0] == pb.offsets[k][0]                          # one CUDA graph serves both
    pc = PackedBatch.from_batch(b, pin=False, classes=[0, 1])
    assert pc.views(pc.buf)['y_class'].dtype == torch.int64
    # compact feeder record: uint16 graph-local edge ids, no edge attributes
    cp = PackedBatch.from_batch(b, pin=False, idx16=True, edge_attr=False, compact=False)
    cv = cp.views(cp.buf)
    assert cp.nbytes < pb.nbytes - 4 * b.edge_index.size(1) and cv['edge_attr'] is None
    assert cv['edge_index'].dtype == torch.int16 and int(cv['edge_index'].max()) < 200
    first = torch.repeat_interleave(b._node_ptr[:-1].long(), (b._edge_ptr[1:] - b._edge_ptr[:-1]).long())
    assert torch.equal(cv['edge_index'].long() + first, b.edge_index)
    assert torch.equal(cv['x'], b.x) and torch.equal(cv['cluster0'].long(), b.cluster0) and torch.equal(cv['y'], b.y)
    assert cp.layout_key() != pb.layout_key()
    dev_like = torch.zeros(cp.capacity_numel)          # staging-buffer sized views keep the same offsets
    dev_like[:cp.numel] = cp.buf
    assert torch.equal(cp.views(dev_like)['edge_index'], cv['edge_index'])
    # ... and the default compact form: the loader stores every edge in both directions (first half i -> j, second
    # half j -> i, DataSet.py:266-269), so only the first half of every graph's list travels; cluster ids as uint16
    hp = PackedBatch.from_batch(b, pin=False, idx16=True, edge_attr=False)
    assert hp.compact and not cp.compact and hp.layout_key() != cp.layout_key()
    E, N, L1 = b.edge_index.size(1), b.x.size(0), b.cluster1.numel()
    assert cp.nbytes - hp.nbytes >= 4 * (E // 2 + N // 2 + L1 // 2) - 48
    hv = hp.views(hp.buf)
    assert hv['edge_index'].shape == (2, E // 2) and hv['cluster0'].dtype == torch.int16
    local = b.edge_index - first
    for g in range(5):                                                          # graph g: pairs [500 g, 500 g + 500)
        half = hv['edge_index'][:, 500 * g:500 * (g + 1)].long()
        assert torch.equal(half, local[:, 1000 * g:1000 * g + 500])
        assert torch.equal(half.flip(0), local[:, 1000 * g + 500:1000 * (g + 1)])
    assert torch.equal(hv['cluster0'].long(), b.cluster0) and torch.equal(hv['cluster1'].long(), b.cluster1)
    assert torch.equal(hv['x'], b.x) and torch.equal(hv['node_ptr'], b._node_ptr) and torch.equal(hv['edge_ptr'], b._edge_ptr)
    dev_like = torch.zeros(hp.capacity_numel)
    dev_like[:hp.numel] = hp.buf
    assert hp.views(dev_like, capacity=True)['cluster1'].numel() == N
    assert torch.equal(hp.views(dev_like, capacity=True)['cluster1'][:L1], hv['cluster1'])
    # a batch whose edge list is not two mirrored halves keeps the full list
    gs = synthetic.make_graphs('cfg2', count=2, seed=4)
    gs[1].edge_index = gs[1].edge_index[:, torch.randperm(gs[1].edge_index.size(1), generator=torch.Generator().manual_seed(1))]
    assert not PackedBatch.from_batch(Batch.from_data_list(gs), pin=False, idx16=True, edge_attr=False).compact

    class DS(object):
        def len(self):
            return len(graphs)

        def get(self, i):
            return graphs[i]
    sizes = [bb.num_graphs for bb in DataLoader(DS(), batch_size=2)]
    assert sizes == [2, 2, 1]


def test_synthetic_graphs_follow_the_fixture_shape():
    from deeprank_gnn_b200 import synthetic
    for g in synthetic.make_graphs('cfg2', count=6, seed=0):
        n = g.x.size(0)
        row, col = g.edge_index
        e = row.numel() // 2
        assert n == 200 and 2 * e == 1000
        assert torch.equal(row[:e], col[e:]) and torch.equal(col[:e], row[e:])
        assert (torch.bincount(row, minlength=n) > 0).all()
        assert (g.edge_index[0] != g.edge_index[1]).all()
        assert torch.unique(row * n + col).numel() == 2 * e
        assert g.cluster1.numel() == g.cluster0.unique().numel()
        na = (n + 1) // 2
        assert bool(((row < na) != (col < na)).all())                           # strictly inter-chain
    assert synthetic.make_graph(200, 1000, 32, 5).x.equal(synthetic.make_graph(200, 1000, 32, 5).x)


def test_flat_parameter_layout_and_reference_names():
    from deeprank_gnn_b200.engine import NetSpec
    from oracle import nets as onets
    for kind, cls in (('GINet', onets.GINet), ('sGAT', onets.sGAT), ('FoutNet', onets.FoutNet)):
        for F, out, hidden in ((3, 1, (16, 32)), (48, 2, (16, 32)), (32, 1, (32, 64))):
            spec = NetSpec(kind, F, out, 1, hidden)
            ref = cls(F, out, 1, hidden=hidden).state_dict()
            assert spec.reference_order() == list(ref.keys())
            shapes = {n: tuple(s) for n, s, _ in spec.param_shapes()}
            assert shapes == {k: tuple(v.shape) for k, v in ref.items()}


def test_shard_indices_balance_and_cover():
    from deeprank_gnn_b200.parallel import shard_indices
    rng = np.random.default_rng(0)
    costs = [int(c) for c in rng.integers(400, 9000, size=512)]
    parts = shard_indices(costs, 8, balance=True)
    assert sorted(i for p in parts for i in p) == list(range(512))
    assert all(len(p) == 64 for p in parts)
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) / (sum(loads) / 8) < 1.02
    flat = shard_indices(list(range(10)), 4, balance=False)
    assert flat == [[0, 1, 2], [3, 4, 5], [6, 7], [8, 9]]


WORKER = r'''
import os, sys, copy
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import torch, torch.distributed as dist
from deeprank_gnn_b200 import synthetic, parallel
from helpers import to_oracle_batch
from oracle import nets as onets, step as ostep
rank, world, _ = parallel.init_distributed('gloo')
graphs = synthetic.make_graphs(dict(nodes=(30, 120), edges_per_node=6, feat=8, batch=10), count=10, seed=7)
torch.manual_seed(0)
model = onets.sGAT(8, 1, 1).eval()
# full-batch gradient (what one process computes)
full = copy.deepcopy(model)
pred = full(to_oracle_batch(graphs)).reshape(-1)
y = torch.cat([g.y for g in graphs])
torch.nn.MSELoss()(pred, y).backward()
ref = torch.cat([p.grad.reshape(-1) for p in full.parameters()])
# sharded: local sum / B_global, then ONE all-reduce of the flat gradient buffer
mine = parallel.shard_graphs(graphs, world, rank, balance=True)
pred = model(to_oracle_batch(mine)).reshape(-1)
y = torch.cat([g.y for g in mine])
(((pred - y) ** 2).sum() / len(graphs)).backward()
flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
parallel.all_reduce_flat_(flat)
err = float((flat - ref).abs().max())
sizes = [None] * world
dist.all_gather_object(sizes, len(mine))
assert sum(sizes) == 10 and max(sizes) - min(sizes) <= 1, sizes
assert err < 1e-5, err
print('rank', rank, 'ok', err)
dist.destroy_process_group()
'''


def test_two_rank_gloo_gradient_allreduce_equals_full_batch(tmp_path):
    """world_size 2 on CPU (gloo): graph sharding + loss = sum/B_global + one flat all-reduce
    reproduces the single-process gradient of MSELoss(mean)."""
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29631', str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count('ok') == 2


def test_structure_blob_layout_macros_match_the_kernels():
    """The blob addressing used by tests / host code (48 g + 12 n0 + 4 e0 words, 9 n + 5 + 3 m payload) is what the
    header declares."""
    hdr = open(os.path.join(ROOT, 'include', 'drgnn.h')).read()
    assert '#define DRGNN_BLOB_HEADER 32' in hdr
    assert '(48 * (int64_t)(g) + 12 * (int64_t)(n0) + 4 * (int64_t)(e0))' in hdr
    assert '(48 * (int64_t)(B) + 12 * (int64_t)(N) + 4 * (int64_t)(E) + 16)' in hdr
    from deeprank_gnn_b200 import ops
    ints, _f, _l = ops._structure_sizes(3, 100, 400, 100, 0)
    assert ints['blob'] == 48 * 3 + 12 * 100 + 4 * 400 + 16
    # every graph's payload fits its slot: 32 + 9 n + 5 + 3 m <= 48 + 12 n + 4 m
    for n, m in ((0, 0), (1, 0), (5, 40), (200, 1000)):
        assert ((32 + 9 * n + 5 + 3 * m + 3) & ~3) <= 48 + 12 * n + 4 * m


def test_rotation_chunk_selection():
    """Chunk CUDA graphs of Engine.train_resident: only for rotations whose structure slots cycle cleanly."""
    from deeprank_gnn_b200.engine import rotation_chunk
    assert rotation_chunk([i % 4 for i in range(64)], 4, 2) == 16
    assert rotation_chunk([i % 4 for i in range(8)], 4, 2) == 8
    assert rotation_chunk([i % 4 for i in range(20)], 4, 2) == 5
    assert rotation_chunk([i % 4 for i in range(4)], 4, 2) == 4
    assert rotation_chunk([i % 4 for i in range(6)], 4, 2) == 0        # slots collide across the wrap-around
    assert rotation_chunk([0, 1], 4, 2) == 0                            # fewer batches than slots
    assert rotation_chunk([0, 0, 1, 1, 2, 2, 3, 3], 4, 2) == 0          # neighbours share a slot
    assert rotation_chunk([i % 4 for i in range(28)], 4, 2) == 4        # 28 = 4 * 7


def test_packed_cache_roundtrip(tmp_path):
    """Packed on-disk cache of feeder records: what comes back from the memory map is the
    record that went in - sections, layout key, bounds, molecule names - for compact and plain records."""
    import warnings
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.data import Batch, PackedBatch, PackedCache
    from deeprank_gnn_b200.DataSet import HDF5DataSet
    from conftest import FIXTURE
    ds = HDF5DataSet(database=FIXTURE, node_feature=['type', 'polarity', 'bsa'], target='irmsd')
    fix = [ds.get(i) for i in range(ds.len())]
    packed = [PackedBatch.from_batch(Batch.from_data_list(fix[:4]), pin=False, idx16=True, edge_attr=False),
              PackedBatch.from_batch(Batch.from_data_list(fix[4:]), pin=False),
              PackedBatch.from_batch(Batch.from_data_list(synthetic.make_graphs('cfg2', count=3, seed=1)), pin=False,
                                     idx16=True, classes=[0, 1])]
    path = str(tmp_path / 'epoch.drgnnpc')
    assert PackedCache.build(path, packed) == 3
    assert os.path.getsize(path) % 64 == 0
    cache = PackedCache(path)
    assert len(cache) == 3
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')          # torch warns about the read-only memory map
        for src, got in zip(packed, cache):
            assert got.layout_key() == src.layout_key() and got.numel == src.numel and got.has_y == src.has_y
            assert got.mol == (list(src.mol) if src.mol is not None else None)
            assert torch.equal(got.buf, src.buf)
            va, vb = src.views(src.buf), got.views(got.buf)
            for k in ('x', 'y', 'edge_index', 'cluster0', 'cluster1', 'node_ptr', 'edge_ptr', 'c1_ptr'):
                assert torch.equal(va[k], vb[k]), k
        again = PackedCache(path)[1]
        assert torch.equal(again.views(again.buf)['edge_attr'], packed[1].views(packed[1].buf)['edge_attr'])
    with open(str(tmp_path / 'junk'), 'wb') as f:
        f.write(b'not a cache, definitely')
    with pytest.raises(ValueError):
        PackedCache(str(tmp_path / 'junk'))


def test_packed_batch_with_almost_as_many_clusters_as_nodes():
    """a batch whose number of level-0 clusters falls into the same 4-word bucket as its node
    count (pad4(L1) == pad4(N), L1 < N) must pack: only device staging buffers get the capacity-sized
    ``cluster1`` view (``views(buf, capacity=True)``)."""
    from deeprank_gnn_b200.data import Batch, Data, PackedBatch
    g = Data(x=torch.randn(4, 3), edge_index=torch.tensor([[0, 1, 2, 3], [1, 0, 3, 2]]),
             edge_attr=torch.rand(4, 1), y=torch.tensor([0.5]), pos=torch.randn(4, 3),
             cluster0=torch.tensor([0, 1, 2, 2]), cluster1=torch.tensor([0, 0, 1]))
    b = Batch.from_data_list([g])
    pb = PackedBatch.from_batch(b, pin=False)
    assert pb.L1 == 3 and pb.N == 4 and pb.numel == pb.capacity_numel
    v = pb.views(pb.buf)
    assert v['cluster1'].numel() == 3 and v['cluster1'].tolist() == [0, 0, 1]
    cap = pb.views(torch.zeros(pb.capacity_numel), capacity=True)
    assert cap['cluster1'].numel() == 4
    with pytest.raises(ValueError):
        pb.views(torch.zeros(pb.capacity_numel - 4), capacity=True)


def test_nccl_bridge_binds_at_run_time_and_validates_arguments(lib):
    """drgnn_nccl_* (csrc/nccl_bridge.cu): libdrgnn.so carries no link-time NCCL dependency, binds the library
    with dlopen and refuses malformed calls with an error code + message (no compute, no GPU)."""
    import ctypes as C
    out = subprocess.run(['readelf', '-d', lib._name], capture_output=True, text=True).stdout
    assert 'libnccl' not in out                                  # bound with dlopen, not linked
    assert lib.drgnn_nccl_allreduce(None, None, 4, None) != 0
    assert b'no communicator' in lib.drgnn_last_error()
    assert lib.drgnn_nccl_destroy(None) == 0                     # closing nothing is fine
    if not lib.drgnn_nccl_available():
        pytest.skip('no NCCL library on this host')
    from deeprank_gnn_b200.parallel import NcclComm
    a, b = NcclComm.unique_id(), NcclComm.unique_id()
    assert len(a) == len(b) == 128 and a != b
    comm = C.c_void_p()
    buf = C.create_string_buffer(a, 128)
    assert lib.drgnn_nccl_init(C.byref(comm), 2, 5, C.cast(buf, C.c_void_p)) != 0
    assert b'outside a world of 2' in lib.drgnn_last_error()
    assert lib.drgnn_nccl_init(None, 1, 0, C.cast(buf, C.c_void_p)) != 0


def test_named_network_entry_points_check_their_kind(lib):
    """drgnn_sgat_step / drgnn_fout_step (SURVEY 8b names) are the general cluster launch bound to ONE network:
    a record of another kind is refused before anything is launched."""
    import ctypes as C
    from deeprank_gnn_b200 import _lib
    s = _lib.NetStepArgs()
    for name, kind in (('drgnn_sgat_step', 1), ('drgnn_fout_step', 2)):
        for k in (0, 1, 2):
            if k == kind:
                continue
            s.kind = k
            assert getattr(lib, name)(C.byref(s), None) == -1
            assert b'must carry kind %d' % kind in lib.drgnn_last_error()
        assert getattr(lib, name)(None, None) == -1


def test_ctypes_struct_offsets_equal_the_c_compilers(lib, tmp_path):
    """Every argument record of the C-ABI: sizeof and the offset of EVERY field as gcc lays the header's struct out
    equal the ctypes mirror's (names and order alone would miss an int32 / int64 / pointer mix-up, which shifts every
    later pointer).  The header is compiled as plain C99 - it is a C interface."""
    import ctypes
    from deeprank_gnn_b200 import _lib
    pairs = [('drgnn_structure_io', _lib.StructureIO), ('drgnn_aggregate_args', _lib.AggregateArgs),
             ('drgnn_linear_args', _lib.LinearArgs), ('drgnn_linear_wgrad_args', _lib.LinearWgradArgs),
             ('drgnn_head_args', _lib.HeadArgs), ('drgnn_ginet_fused_args', _lib.GinetFusedArgs),
             ('drgnn_ginet_step_args', _lib.GinetStepArgs), ('drgnn_net_step_args', _lib.NetStepArgs),
             ('drgnn_peer_comm', _lib.PeerComm), ('drgnn_peer_adam_args', _lib.PeerAdamArgs),
             ('drgnn_feed_step', _lib.FeedStep)]
    src = ['#include <stddef.h>', '#include <stdio.h>', '#include "drgnn.h"', 'int main(void) {']
    for cname, cls in pairs:
        src.append('  printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _t in cls._fields_:
            src.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    src += ['  return 0;', '}']
    c = tmp_path / 'layout.c'
    c.write_text('\n'.join(src))
    exe = tmp_path / 'layout'
    r = subprocess.run(['gcc', '-std=c99', '-Wall', '-Werror', '-pedantic', '-I', os.path.join(ROOT, 'include'), str(c), '-o', str(exe)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs:
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _t in cls._fields_:
            assert int(got['%s.%s' % (cname, fname)]) == getattr(cls, fname).offset, '%s.%s' % (cname, fname)


def test_ctypes_prototypes_equal_the_headers(lib):
    """Every entry point: number and width class of the arguments (pointer / int32 / int64 / float) and the result
    type declared in include/drgnn.h equal the ctypes prototype (_lib._SIGNATURES) - a call through a prototype with
    a narrower integer or a missing argument would pass garbage in a register."""
    import ctypes as C
    from deeprank_gnn_b200 import _lib
    hdr = open(os.path.join(ROOT, 'include', 'drgnn.h')).read()
    hdr = re.sub(r'/\*.*?\*/', '', hdr, flags=re.S)
    protos = re.findall(r'^\s*((?:const\s+)?[A-Za-z0-9_]+\s*\*?)\s*(drgnn_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', hdr, flags=re.M)
    assert len(protos) == len(_lib._SIGNATURES)

    def klass_c(decl):
        decl = decl.strip()
        if '*' in decl or '[' in decl:
            return 'ptr'
        base = re.sub(r'\b(const|unsigned|signed)\b', '', decl).split()[0]
        return {'int32_t': 'i32', 'int': 'i32', 'uint32_t': 'i32', 'int64_t': 'i64', 'uint64_t': 'i64', 'float': 'f32'}[base]

    def klass_py(t):
        if t in (C.c_void_p, C.c_char_p) or hasattr(t, 'contents'):
            return 'ptr'
        return {C.c_int32: 'i32', C.c_int: 'i32', C.c_uint32: 'i32', C.c_int64: 'i64', C.c_uint64: 'i64', C.c_float: 'f32'}[t]

    for ret, name, args in protos:
        res, argtypes = _lib._SIGNATURES[name]
        args = args.strip()
        want = [] if args in ('', 'void') else [klass_c(a) for a in args.split(',')]
        assert [klass_py(t) for t in argtypes] == want, name
        assert klass_py(res) == klass_c(ret), name


def test_library_carries_blackwell_native_sass(lib):
    """The built library is sm_100a code with the Blackwell-only machinery the design claims (cuobjdump, no GPU):
    tcgen05 MMA + TMEM loads + tcgen05 commit in the dense transform (UTCHMMA, LDTM, UTCBAR), TMA bulk copies and
    mbarriers in the step / structure / aggregation kernels (UBLKCP, SYNCS), thread-block-cluster barriers
    (UCGABAR) - and no kernel compiled for another architecture."""
    import shutil
    if not shutil.which('cuobjdump'):
        pytest.skip('cuobjdump not on PATH')
    from deeprank_gnn_b200 import _lib
    elfs = subprocess.run(['cuobjdump', '-lelf', _lib.LIB_PATH], capture_output=True, text=True).stdout.split()
    cubins = [e for e in elfs if e.endswith('.cubin')]
    kernels = [e for e in cubins if not e.startswith('libdrgnn.')]          # (the link step's empty stub aside)
    assert len(kernels) >= 12 and all('.sm_100a.' in e for e in kernels), cubins
    sass = subprocess.run(['cuobjdump', '-sass', _lib.LIB_PATH], capture_output=True, text=True).stdout
    for mnemonic, least in (('UTCHMMA', 3), ('LDTM', 1), ('UTCBAR', 1), ('UBLKCP', 20), ('SYNCS', 100), ('UCGABAR', 8)):
        assert sass.count(mnemonic) >= least, (mnemonic, sass.count(mnemonic))
