"""Fused training / scoring engine for the reference networks (GINet, sGAT, FoutNet).

It replaces the body of the reference's per-batch loop (``NeuralNet._epoch``,
``deeprank_gnn/NeuralNet.py:490-503``: zero_grad, ``model(batch)``, ``format_output``,
loss, ``backward``, ``Adam.step``) by a fixed sequence of C-ABI kernel launches over flat,
pre-allocated buffers:

    structure pass (2 launches, integer)        get_preloaded_cluster / consecutive_cluster /
                                                 pool_edge / pool_batch   community_pooling.py:25-30,197-224
    conv1 = aggregate -> transform(+ReLU)       ginet.py:50-73 | sGAT.py:62-93 | foutnet.py:56-82
    cluster max-pool (level 0)                  community_pooling.py:201
    conv2 = aggregate -> transform(+ReLU)       on the coarsened graph
    cluster max-pool (level 1)                  max_pool_x (ginet.py:114,129)
    graph mean read-out, fc1(+ReLU,+dropout), fc2   ginet.py:133-139
    loss + dLoss/dpred, hand-written backward, [NCCL all-reduce], flat Adam

Both GINet branches run in the same launches (their graphs are identical):
conv1 shares the aggregated input and concatenates the two weights, conv2 is a 2-group
transform.  No host synchronisation happens inside a step, so the launch sequence can be
captured in a CUDA graph (``graph=True``) and replayed.

Parameters live in one flat fp32 buffer; ``state_dict()`` / ``load_state_dict()`` use the
reference's names and shapes, so reference checkpoints load unchanged.
"""
import os
import time
from collections import OrderedDict

import torch

from . import ops
from ._lib import DrgnnError

F32, I32, I64 = torch.float32, torch.int32, torch.int64


def rotation_chunk(sslots, n_slots, lookahead):
    """Steps per chunk CUDA graph for a rotation over batches whose structure slots are ``sslots`` (0: chunk
    graphs do not apply).  Needs a chunk length that divides the rotation, and slots such that a batch, the
    ``n_slots - 1`` batches before it never collide while the batch ``n_slots`` before it uses the same slot -
    then a structure pass ``lookahead`` steps ahead only has to wait for the step that read its slot last."""
    R = len(sslots)
    if R < n_slots:
        return 0
    for i in range(R):
        s_i = sslots[i]
        if any(sslots[(i - k) % R] == s_i for k in range(1, n_slots)) or sslots[(i - n_slots) % R] != s_i:
            return 0
    for C in (16, 12, 8, 6, 5, 4):
        if R % C == 0 and C > lookahead:
            return C
    return 0


class NetSpec(object):
    """Architecture description of one of the reference networks."""

    def __init__(self, kind, input_shape, output_shape=1, input_shape_edge=1, hidden=(16, 32), dropout=None, layers=2):
        kind = {'GINet': 'ginet', 'sGAT': 'sgat', 'FoutNet': 'fout'}.get(kind, kind).lower()
        # layers = 3 (sGAT / FoutNet): a third conv layer h2 -> h2 on the coarsened graph before the level-1 max-pool -
        # the "sGAT 3-layer" throughput variant of BASELINE config 3 (the reference nets have two layers)
        self.layers = int(layers)
        if self.layers not in (2, 3) or (self.layers == 3 and kind == 'ginet'):
            raise ValueError('layers must be 2, or 3 for sGAT / FoutNet')
        if kind not in ('ginet', 'sgat', 'fout'):
            raise ValueError('unknown network %r (GINet, sGAT or FoutNet)' % kind)
        self.kind = kind
        self.F = int(input_shape)
        self.out = int(output_shape)
        self.ne = int(input_shape_edge) if input_shape_edge else 1
        self.h1, self.h2 = int(hidden[0]), int(hidden[1])
        if self.h1 % 4 or self.h2 % 4:
            raise ValueError('hidden widths must be multiples of 4')
        self.nb = 2 if kind == 'ginet' else 1                 # branches
        self.C1, self.C2 = self.nb * self.h1, self.nb * self.h2
        self.Kin1 = self.F if kind == 'ginet' else 2 * self.F
        self.Kin2 = self.C1 if kind == 'ginet' else 2 * self.h1
        self.Hd = 4 * self.h2 if kind == 'ginet' else 2 * self.h2     # fc1 width (ginet.py:94, sGAT.py:109)
        self.dropout = (0.4 if kind == 'ginet' else 0.0) if dropout is None else float(dropout)   # ginet.py:97
        self.w_layout = 0 if kind == 'ginet' else 1

    def param_shapes(self):
        """Flat layout: (reference state_dict name, shape, live?) in storage order."""
        F, h1, h2, out, ne, Hd = self.F, self.h1, self.h2, self.out, self.ne, self.Hd
        if self.kind == 'ginet':
            p = [('conv1.fc.weight', (h1, F), True), ('conv1_ext.fc.weight', (h1, F), True),
                 ('conv2.fc.weight', (h2, h1), True), ('conv2_ext.fc.weight', (h2, h1), True),
                 ('fc1.weight', (Hd, 2 * h2), True), ('fc1.bias', (Hd,), True),
                 ('fc2.weight', (out, Hd), True), ('fc2.bias', (out,), True)]
            # parameters of the dead attention path (alpha == 1): zero gradient
            for name, cin, cout in (('conv1', F, h1), ('conv2', h1, h2), ('conv1_ext', F, h1), ('conv2_ext', h1, h2)):
                p.append((name + '.fc_edge_attr.weight', (ne, ne), False))
                p.append((name + '.fc_attention.weight', (1, 2 * cout + ne), False))
        elif self.kind == 'sgat':
            p = [('conv1.weight', (2 * F, h1), True), ('conv1.bias', (h1,), True),
                 ('conv2.weight', (2 * h1, h2), True), ('conv2.bias', (h2,), True)]
            if self.layers == 3:
                p += [('conv3.weight', (2 * h2, h2), True), ('conv3.bias', (h2,), True)]
            p += [('fc1.weight', (Hd, h2), True), ('fc1.bias', (Hd,), True),
                  ('fc2.weight', (out, Hd), True), ('fc2.bias', (out,), True)]
        else:
            p = [('conv1.Wc', (F, h1), True), ('conv1.Wn', (F, h1), True), ('conv1.bias', (h1,), True),
                 ('conv2.Wc', (h1, h2), True), ('conv2.Wn', (h1, h2), True), ('conv2.bias', (h2,), True)]
            if self.layers == 3:
                p += [('conv3.Wc', (h2, h2), True), ('conv3.Wn', (h2, h2), True), ('conv3.bias', (h2,), True)]
            p += [('fc1.weight', (Hd, h2), True), ('fc1.bias', (Hd,), True),
                  ('fc2.weight', (out, Hd), True), ('fc2.bias', (out,), True)]
        return p

    # names in the order the reference's nn.Module registers them (state_dict order)
    def reference_order(self):
        if self.kind == 'ginet':
            names = []
            for c in ('conv1', 'conv2', 'conv1_ext', 'conv2_ext'):
                names += [c + '.fc.weight', c + '.fc_edge_attr.weight', c + '.fc_attention.weight']
            return names + ['fc1.weight', 'fc1.bias', 'fc2.weight', 'fc2.bias']
        return [name for name, _shape, _live in self.param_shapes()]


def _pad4(n):
    return (n + 3) & ~3


class FlatParams(object):
    """One fp32 buffer holding every parameter, with named views."""

    def __init__(self, spec, device):
        self.spec = spec
        self.slots = OrderedDict()
        o = 0
        for name, shape, live in spec.param_shapes():
            n = 1
            for s in shape:
                n *= s
            self.slots[name] = (o, n, shape, live)
            # conv weight pairs (GINet branches, Fout Wc/Wn) are consumed as ONE matrix and must stay
            # adjacent: their sizes are multiples of 4 (hidden widths are), so the padding is a no-op
            if name.split('.')[0].startswith('conv') and live and len(shape) == 2:
                assert n % 4 == 0
            o += _pad4(n)
        self.numel = _pad4(o)
        self.data = torch.zeros(self.numel, dtype=F32, device=device)

    def view(self, buf, name):
        o, n, shape, _ = self.slots[name]
        return buf[o:o + n].view(shape)

    def offset(self, name):
        return self.slots[name][0]


class Workspace(object):
    """Activation / gradient buffers for batches up to (B, N, E) - allocated once, reused."""

    def __init__(self, spec, B, N, E, device, n_params=0):
        s = spec
        self.B, self.N, self.E = B, N, E
        z = lambda *shape: torch.zeros(*shape, dtype=F32, device=device)
        self.Zin1 = z(N, s.Kin1)
        self.Z1 = z(N, s.C1)
        self.P1 = z(N, s.C1)
        self.arg0 = torch.zeros(N, s.C1, dtype=I32, device=device)
        self.Zin2 = z(N, s.Kin2)
        self.Z2 = z(N, s.C2)
        self.P2 = z(N, s.C2)
        self.arg1 = torch.zeros(N, s.C2, dtype=I32, device=device)
        self.R = z(B, s.C2)
        self.H = z(B, s.Hd)
        self.pred = z(B, s.out)
        self.keep = z(B, s.Hd)
        self.loss = z(1)
        # per-row scalars of the mean aggregations (sGAT / Fout)
        self.s0, self.post0, self.s1, self.post1 = z(N), z(N), z(N), z(N)
        # gradients
        self.dpred = z(B, s.out)
        self.dH = z(B, s.Hd)
        self.dR = z(B, s.C2)
        self.dP2 = z(N, s.C2)
        self.dZ2 = z(N, s.C2)
        self.dZin2 = z(N, s.Kin2)
        self.dP1 = z(N, s.C1)
        self.dZ1 = z(N, s.C1)
        need = max(ops.linear_wgrad_work_floats(N, s.Kin1, s.C1, 1),
                   ops.linear_wgrad_work_floats(N, s.Kin2 // (s.nb if s.kind == 'ginet' else 1), s.h2, s.nb),
                   ops.linear_wgrad_work_floats(B, s.C2, s.Hd, 1),
                   ops.linear_wgrad_work_floats(B, s.Hd, s.out, 1))
        self.wwork = z(need)
        # per-graph weight-gradient partials of the fused per-graph backward (GINet)
        self.partial = z(B, s.C1 * s.F + s.nb * s.h2 * s.h1) if s.kind == 'ginet' else None
        # per-graph rows shaped like the flat gradient buffer (+ the loss slot) for the whole-step kernel
        self.partial_full = z(B, n_params + 4)


class DeviceBatch(object):
    """The tensors of one mini-batch the engine consumes, resident on the device."""
    __slots__ = ('x', 'edge_index', 'edge_attr', 'cluster0', 'cluster1', 'node_ptr', 'edge_ptr', 'c1_ptr', 'y',
                 'y_class', 'B', 'N', 'E', 'L1', 'L1b', 'max_n', 'max_e', 'max_k0', 'max_k1', 'mol', 'key', 'sslot',
                 'edge_half')

    @staticmethod
    def from_batch(batch, device, classes=None):
        """From a collated ``Batch`` (``data.Batch.from_data_list``), reference dtypes (int64
        indices).  ``classes``: class list for classification targets (NeuralNet.py:616-631)."""
        d = DeviceBatch()
        if getattr(batch, '_node_ptr', None) is None or getattr(batch, '_edge_ptr', None) is None:
            raise DrgnnError('the engine needs a Batch collated by Batch.from_data_list (graph pointers)')
        if getattr(batch, 'cluster0', None) is None or getattr(batch, 'cluster1', None) is None:
            raise DrgnnError('the batch carries no cluster0 / cluster1 (run PreCluster first, DataSet.py:45-88)')
        mv = lambda t: None if t is None else t.to(device, non_blocking=True)
        x = batch.x if batch.x.dim() == 2 else batch.x.unsqueeze(-1)
        d.x = mv(x).float().contiguous()
        d.edge_index = mv(batch.edge_index).contiguous()
        ea = getattr(batch, 'edge_attr', None)
        if ea is not None:
            ea = mv(ea).float()
            ea = ea.unsqueeze(-1) if ea.dim() == 1 else ea
            ea = ea.contiguous()
        d.edge_attr = ea
        d.cluster0, d.cluster1 = mv(batch.cluster0).contiguous(), mv(batch.cluster1).contiguous()
        d.node_ptr, d.edge_ptr = mv(batch._node_ptr), mv(batch._edge_ptr)
        d.c1_ptr = mv(batch._c1_ptr)
        y = getattr(batch, 'y', None)
        d.y = None if y is None else mv(y.float().reshape(-1).contiguous())
        d.y_class = None
        if classes is not None and y is not None:
            c2i = {int(c): i for i, c in enumerate(classes)}
            d.y_class = mv(torch.tensor([c2i[int(t)] for t in y.reshape(-1).tolist()], dtype=I64))
        d.B, d.N, d.E, d.L1 = batch.num_graphs, d.x.size(0), d.edge_index.size(1), d.cluster1.numel()
        d.max_n, d.max_e = int(batch._max_n), int(batch._max_e)
        d.max_k0, d.max_k1 = getattr(batch, '_max_k0', None), getattr(batch, '_max_k1', None)
        d.mol = getattr(batch, 'mol', None)
        d.key = None
        d.sslot = 0
        d.edge_half = False
        d.L1b = d.L1                     # upper bound of level-1 rows used for launch sizes
        return d

    @staticmethod
    def from_packed(pb, dev_buf):
        """Views into ``dev_buf`` (device float32 buffer holding a copy of ``pb.buf``)."""
        d = DeviceBatch()
        v = pb.views(dev_buf, capacity=True)
        d.x, d.edge_attr, d.y, d.y_class = v['x'], v['edge_attr'], (v['y'] if pb.has_y else None), v['y_class']
        d.edge_index, d.cluster0, d.cluster1 = v['edge_index'], v['cluster0'], v['cluster1']
        d.node_ptr, d.edge_ptr, d.c1_ptr = v['node_ptr'], v['edge_ptr'], v['c1_ptr']
        d.B, d.N, d.E, d.L1, d.max_n, d.max_e = pb.B, pb.N, pb.E, pb.L1, pb.max_n, pb.max_e
        d.max_k0, d.max_k1 = pb.max_k0, pb.max_k1
        d.mol = pb.mol
        d.key = pb.layout_key()
        d.sslot = 0
        d.edge_half = bool(pb.compact)   # edge_index is [2, E/2]: the first half of every graph's mirrored edge list
        d.L1b = pb.N                     # fixed bound: batches of one layout replay the same CUDA graph
        return d


class Engine(object):
    STRUCT_SLOTS = 4      # structure passes may run this many batches ahead of the step (weight-independent work)

    def __init__(self, net, input_shape, output_shape=1, input_shape_edge=1, hidden=(16, 32), device='cuda',
                 task='reg', class_weights=None, transform_sigmoid=False, lr=0.01, betas=(0.9, 0.999), eps=1e-8,
                 dropout=None, graph=False, tiled=True, fused_head=True, fused_graph=True, process_group=None, seed=None,
                 peer_comm=True, layers=2):
        self.spec = NetSpec(net, input_shape, output_shape, input_shape_edge, hidden, dropout, layers)
        self.device = torch.device(device)
        if self.device.type != 'cuda':
            raise DrgnnError('the engine runs on a CUDA device only (no CPU fallback)')
        self.task = task
        if task not in ('reg', 'class'):
            raise ValueError("task must be 'reg' or 'class'")
        self.transform_sigmoid = bool(transform_sigmoid)
        self.class_weights = None if class_weights is None else \
            torch.as_tensor(class_weights, dtype=F32).to(self.device).contiguous()
        self.lr, self.betas, self.eps = float(lr), betas, float(eps)
        self.training = True
        self.use_graph = bool(graph)
        self.tiled = tiled             # True: per-graph TMA kernel for large batches; 'force': always; False: never
        self.pg = process_group
        dist = torch.distributed
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        self.comm = None                 # parallel.PeerComm: fused exchange + Adam over NVLink peer memory
        self.comm_error = None
        self._no_exchange = False        # CUDA-graph warm-up: run the kernels without talking to the peers
        self._want_adam = True
        self._reduced = False
        self.params = FlatParams(self.spec, self.device)
        # gradients and the loss share one buffer so that a multi-GPU step needs ONE all-reduce
        self._grads_full = torch.zeros(self.params.numel + 4, dtype=F32, device=self.device)
        self.grads = self._grads_full[:self.params.numel]
        self._loss_slot = self._grads_full[self.params.numel:self.params.numel + 1]
        self.exp_avg = torch.zeros_like(self.params.data)
        self.exp_avg_sq = torch.zeros_like(self.params.data)
        self.step_dev = torch.zeros(4, dtype=F32, device=self.device)      # [0] Adam step count, [1] ticket of the fused optimiser
        self.ws = None
        self.structs = [None] * self.STRUCT_SLOTS     # structure slots: passes of the next batches overlap step i
        self._last_struct = None
        self._graphs = {}
        self._staging = {}
        self._dbatch = {}
        self._copy_stream = None
        from . import _lib
        self._sms = int(_lib.load().drgnn_device_sms())
        self.fused_head = bool(fused_head)
        self.fused_graph = bool(fused_graph)     # per-graph fused GINet forward / backward kernels
        self._fused_fit = {}
        self._graph_done = False
        self._all_done = False
        self._all_done_kernel = False    # the last training step ran a whole-step kernel
        self._adam_done = False
        self.fuse_adam = True
        self.step_variant = int(os.environ.get('DRGNN_STEP_VARIANT', '0'))   # 0 pick, 1 single-CTA kernel, 2 cluster kernel
        self.native_feed = os.environ.get('DRGNN_NATIVE_FEED', '1') != '0'   # train_batches loop issued from C
        self._feed_keep = None
        self.feed_issue_us = None
        self._read_stream = None
        self._read_ring = None
        self._host_out = None
        self._primed = None              # (rotation, index): structure passes train_resident left done
        self.fuse_comm = os.environ.get('DRGNN_FUSE_COMM', '1') != '0'   # peer exchange inside the step kernel
        self._cur_B_global = None
        self._last_exchange = None
        self.rotation_graph = os.environ.get('DRGNN_ROTATION_GRAPH', '1') != '0'   # train_resident: one CUDA graph per rotation
        self.blob_structure = os.environ.get('DRGNN_BLOB_STRUCTURE', '1') != '0'   # one-launch bitmap structure pass
        self.keep_intermediates = False   # cluster kernel: also mirror AX / Z1 / argmax ... to global memory (tests)
        self.fuse_reduce = os.environ.get('DRGNN_FUSE_REDUCE', '1') != '0'   # gradient reduction + Adam behind a grid barrier
        # general cluster whole-step kernel (csrc/fused_step3.cuh): sGAT / FoutNet, and GINet graphs too large for
        # the CTA-pair kernel (node dimension tiled over the cluster, neighbours through distributed shared memory)
        self.step3 = os.environ.get('DRGNN_STEP3', '1') != '0'
        self.step3_tiles = int(os.environ.get('DRGNN_STEP3_TILES', '0'))      # 0: smallest tile count that fits
        # dense products of the fused step on the tensor cores (mma.sync 3xTF32, error ~1e-6) instead of FFMA tiles;
        # opt-in: measured slower than the FFMA register tiles at every BASELINE config (profiles/README.md, round 2)
        self.fused_tc = os.environ.get('DRGNN_FUSED_TC', '0') != '0'
        # head v2 of the CTA-pair kernel (flag bit 6 of drgnn_ginet_step: fc2 / loss / dLoss/dpred in every warp, fc1.weight
        # gradient rows formed inside the in-kernel reduction instead of stored and re-read; bit-identical results)
        self.head_v2 = os.environ.get('DRGNN_HEAD_V2', '1') != '0'
        # general cluster kernel on a grid larger than the device: clusters take the graphs largest first
        self.step3_lpt = os.environ.get('DRGNN_STEP3_LPT', '1') != '0'
        self._sm_count = torch.cuda.get_device_properties(self.device).multi_processor_count
        # The blob structure pass also computes conv1's input rows (the first aggregation depends on the batch only)
        # and the step kernels stage them: '1' always, '0' never, 'auto' when the step grid fits the device (the pass
        # runs beside the step; see _pre_agg_on for the measurements)
        self.pre_agg = os.environ.get('DRGNN_PRE_AGG', 'auto')
        self.phase_timers = False     # diagnostic: block 0 of the fused step kernels records its phase clocks
        self._last_path = None
        self.seed = 0x5EED if seed is None else int(seed)
        if self.world > 1:
            # per-rank dropout stream: the in-kernel mask hashes (seed, step, LOCAL graph, unit), so
            # ranks sharing a seed would draw identical masks for their shards
            rank = dist.get_rank(process_group)
            self.seed = (self.seed ^ (0x9E3779B1 * (rank + 1))) & 0xffffffff
        self._head_fits = ops.head_fits(self.spec.C2, self.spec.Hd, self.spec.out)
        self._head_done = False
        self.launches_per_step = 0
        self.reset_parameters(seed)
        if self.world > 1 and peer_comm and os.environ.get('DRGNN_PEER_COMM', '1') != '0':
            self._open_comm()
        # no peer memory (or DRGNN_PEER_COMM=0): ONE NCCL all-reduce of [gradients | loss] per step - through
        # torch.distributed, or with DRGNN_NCCL_NATIVE=1 through the C-ABI (drgnn_nccl_allreduce on a communicator
        # owned by libdrgnn: the binding a host without PyTorch's process group would use)
        self.nccl = None
        if self.world > 1 and self.comm is None and os.environ.get('DRGNN_NCCL_NATIVE', '0') != '0':
            from .parallel import NcclComm
            self.nccl = NcclComm(group=self.pg)

    def _open_comm(self):
        """Map the peers' exchange regions (CUDA IPC).  Symmetric on all ranks: either every rank
        gets a communicator or none does (then the step falls back to ONE NCCL all-reduce)."""
        from .parallel import PeerComm
        try:
            self.comm = PeerComm(self.params.numel + 4, group=self.pg)
        except DrgnnError as e:
            self.comm, self.comm_error = None, str(e)

    def collective(self):
        """Name of the gradient exchange this engine uses (reported by bench.py)."""
        if self.world == 1:
            return 'none'
        if self.comm is None:
            return 'nccl all_reduce' + (' (drgnn_nccl_allreduce)' if self.nccl is not None else '')
        if self._last_exchange == 'in-kernel':
            return 'peer-memory exchange + rank-ordered sum + Adam inside the step kernel (no extra launch)'
        return 'peer-memory exchange fused with reduce+Adam (1 launch)'

    def step_kernel_name(self):
        """Name of the kernel that carries the step of the last batch (bench.py's roofline entry)."""
        if self._last_path == 'step3':
            return 'net_graph_step3_kernel'
        if self.spec.kind == 'ginet' and self._all_done_kernel:
            return {1: 'ginet_graph_step_kernel', 2: 'ginet_graph_step2_kernel'}.get(ops.ginet_step_last_variant(),
                                                                                      'ginet_graph_step_kernel')
        return 'op-level sequence (aggregate_rows_kernel / linear_fma_kernel / ...)'

    # ---------------------------------------------------------------- parameters
    def reset_parameters(self, seed=None):
        """Reference initialisation: U(+-1/sqrt(size)) for the conv parameters (ginet.py:43-48,
        sGAT.py:57-60, foutnet.py:50-54), nn.Linear defaults for the heads."""
        g = torch.Generator().manual_seed(seed) if seed is not None else None
        s = self.spec
        sd = OrderedDict()
        for name, shape, _live in s.param_shapes():
            layer = name.split('.')[0]
            if layer.startswith('conv'):
                cin = s.F if layer.startswith('conv1') else (s.h1 if layer.startswith('conv2') else s.h2)
                size = 2 * cin if s.kind == 'sgat' else cin
                bound = 1.0 / (size ** 0.5)
            else:   # nn.Linear default: kaiming_uniform(a=sqrt(5)) == U(+-1/sqrt(fan_in)), same for bias
                fan_in = (2 * s.h2 if s.kind == 'ginet' else s.h2) if layer == 'fc1' else s.Hd
                bound = 1.0 / (fan_in ** 0.5)
            sd[name] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        self.load_state_dict(sd)

    def state_dict(self):
        out = OrderedDict()
        for name in self.spec.reference_order():
            out[name] = self.params.view(self.params.data, name).detach().clone()
        return out

    def load_state_dict(self, sd, strict=True):
        missing = [n for n in self.params.slots if n not in sd]
        extra = [n for n in sd if n not in self.params.slots]
        if strict and (missing or extra):
            raise KeyError('state_dict mismatch: missing %s, unexpected %s' % (missing, extra))
        for name, t in sd.items():
            if name in self.params.slots:
                v = self.params.view(self.params.data, name)
                if tuple(t.shape) != tuple(v.shape):
                    raise ValueError('shape mismatch for %s: %s vs %s' % (name, tuple(t.shape), tuple(v.shape)))
                v.copy_(t.to(self.device, F32))

    def named_grads(self):
        return OrderedDict((n, self.params.view(self.grads, n)) for n in self.spec.reference_order())

    def optimizer_state_dict(self):
        """torch.optim.Adam-shaped state (NeuralNet.py:776 stores optimizer.state_dict())."""
        state = {}
        for i, name in enumerate(self.spec.reference_order()):
            state[i] = {'step': self.step_dev[0].detach().cpu().clone().reshape(()),
                        'exp_avg': self.params.view(self.exp_avg, name).detach().clone(),
                        'exp_avg_sq': self.params.view(self.exp_avg_sq, name).detach().clone()}
        group = {'lr': self.lr, 'betas': tuple(self.betas), 'eps': self.eps, 'weight_decay': 0, 'amsgrad': False,
                 'params': list(range(len(state)))}
        return {'state': state, 'param_groups': [group]}

    def load_optimizer_state_dict(self, osd):
        names = self.spec.reference_order()
        for i, name in enumerate(names):
            st = osd['state'].get(i)
            if st is None:
                continue
            self.params.view(self.exp_avg, name).copy_(st['exp_avg'].to(self.device, F32))
            self.params.view(self.exp_avg_sq, name).copy_(st['exp_avg_sq'].to(self.device, F32))
            self.step_dev[0] = float(st['step'])
        if osd.get('param_groups'):
            g = osd['param_groups'][0]
            self.lr, self.betas, self.eps = float(g['lr']), tuple(g['betas']), float(g['eps'])

    def train(self, mode=True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    # ---------------------------------------------------------------- buffers
    def _ensure(self, B, N, E):
        """Workspace + structure buffers for batches up to (B, N, E); growing them moves every
        buffer, so captured graphs are dropped."""
        ws = self.ws
        if ws is None or ws.B < B or ws.N < N or ws.E < E:
            nb = max(B, ws.B if ws else 0)
            nn_ = max(N, ws.N if ws else 0)
            ne_ = max(E, ws.E if ws else 0)
            self.ws = Workspace(self.spec, nb, nn_, ne_, self.device, self.params.numel)
            # one buffer [flat gradients | loss, pad | predictions]: a multi-GPU step all-reduces the first
            # numel + 4 floats in ONE call, the host read-back of [loss | predictions] is ONE copy
            n = self.params.numel
            self._grads_full = torch.zeros(n + 4 + nb * self.spec.out, dtype=F32, device=self.device)
            self.grads = self._grads_full[:n]
            self._loss_slot = self._grads_full[n:n + 1]
            self.ws.loss = self._loss_slot
            self.ws.pred = self._grads_full[n + 4:].view(nb, self.spec.out)
            ne_attr = 1 if self.spec.kind == 'sgat' else 0
            self.structs = [ops.Structure(nb, nn_, ne_, nn_, ne_attr, self.device) for _ in range(self.STRUCT_SLOTS)]
            for st in self.structs:
                st.sticky_status = True      # status words are OR-ed over the passes of an epoch, see validate()
            self._graphs.clear()
        return self.ws

    # ---------------------------------------------------------------- structure pass
    def prepare(self, d):
        """Run the structure pass of batch ``d`` into structure slot ``d.sslot`` on the current
        stream.  It depends on the batch only (not on the weights), so callers may run it on a
        side stream while the previous step computes (``train_batches`` / ``train_resident`` do)."""
        self._ensure(d.B, d.N, d.E)
        self._primed = None              # a structure slot changes: train_resident primes its lookahead again
        need_w = self.spec.kind == 'sgat'
        if need_w and d.edge_attr is None:
            raise DrgnnError('sGAT needs edge_attr')
        if need_w and d.edge_attr.size(1) != 1:
            raise DrgnnError('sGAT supports one edge feature (the reference broadcast needs ne in {1, Fout})')
        slot = self.structs[d.sslot]
        if self._blob_only(d):
            # fused whole-step paths: ONE launch writes the per-graph structure blobs (+ edge weights for sGAT)
            # (+ the input rows of conv1's transform: the first aggregation does not depend on the weights)
            pre = self._pre_agg_on(d)
            st = ops.structure_blob(d.node_ptr, d.edge_ptr, d.edge_index, d.cluster0, d.max_n, d.max_e, d.c1_ptr,
                                    d.cluster1, out=slot, L1=d.L1, edge_attr=d.edge_attr if need_w else None,
                                    x=d.x if pre else None, zin_kind=self.spec.kind if pre else None,
                                    edge_half=d.edge_half, max_k=d.max_k0, max_q=d.max_k1)
            assert st is slot               # _ensure sized the slot for this batch
            self._last_struct = st
            return st
        st = ops.structure_build(d.node_ptr, d.edge_ptr, d.edge_index, d.cluster0, d.max_n, d.max_e,
                                 c1_ptr=d.c1_ptr, cluster1=d.cluster1, edge_attr=d.edge_attr if need_w else None,
                                 clusters_are_local=True, mirrors=False, out=slot, L1=d.L1, edge_half=d.edge_half)
        assert st is slot
        self._last_struct = st
        return st

    def _pre_agg_on(self, d):
        """Does the blob structure pass of batch ``d`` also compute conv1's input rows?  (``self.pre_agg``)"""
        mode = self.pre_agg
        if isinstance(mode, bool):
            mode = '1' if mode else '0'
        if mode == '0' or self.spec.F % 4:
            return False
        if mode != 'auto':
            return True
        tiles = self._step3_tiles(d)
        ctas = (tiles if tiles else 1) * self.spec.nb * d.B        # grid of the step kernel
        # measured (driver's command, us per step with / without): cfg2 25.9 / 26.1, cfg3 40.0 / 44.1, cfg4 91.3 / 100.0 -
        # but cfg5 131.2 / 122.2: a step grid larger than the device leaves the longer pass no SM of its own
        return ctas <= self._sm_count

    def _comm_in_kernel(self, d, st):
        """Can the peer-memory exchange run inside the cluster step kernel for batch ``d``?  Decided from
        values that are identical on every rank (equal shards, same device type, same shapes), because
        all ranks must take the same path."""
        s = self.spec
        if not (self.fuse_comm and self.fuse_reduce and st.blob_only and self._cur_B_global is not None
                and d.B * self.world == self._cur_B_global and 2 * d.B <= int(self.comm.struct.max_blocks)):
            return False
        key = (d.max_n, d.max_e, d.max_k0, d.max_k1, 'clusters')
        mc = self._fused_fit.get(key)
        if mc is None:
            smem = ops.ginet_step2_smem_bytes(s.F, s.h1, s.h2, d.max_n, d.max_k0, d.max_k1, d.max_e, s.Hd, s.out)
            mc = ops.ginet_step2_max_clusters(smem) if smem >= 0 else 0
            self._fused_fit[key] = mc
        return d.B <= mc

    def _step3_tiles(self, d):
        """Node tiles of the general cluster step kernel for batch ``d`` (0: that kernel does not apply).
        GINet prefers the tuned CTA-pair kernel whenever a graph fits it."""
        s = self.spec
        if not (self.step3 and self.fused_head and self.fused_graph and d.max_k0 and d.max_k1 and d.cluster1 is not None
                and d.c1_ptr is not None and d.x.size(1) == s.F and s.F % 4 == 0 and self.step_variant == 0):
            return 0
        if s.kind == 'sgat' and (d.edge_attr is None or d.edge_attr.size(1) != 1):
            return 0
        key = (d.max_n, d.max_e, d.max_k0, d.max_k1, 'step3', self.step3_tiles, d.B)
        tiles = self._fused_fit.get(key)
        if tiles is None:
            if s.kind == 'ginet' and ops.ginet_step2_smem_bytes(s.F, s.h1, s.h2, d.max_n, d.max_k0, d.max_k1, d.max_e,
                                                                 s.Hd, s.out) >= 0 and not self.step3_tiles:
                tiles = 0
            elif self.step3_tiles:
                ok = ops.net_step_smem_bytes(s.kind, self.step3_tiles, s.F, s.h1, s.h2, d.max_n, d.max_k0, d.max_k1,
                                             d.max_e, s.Hd, s.out, layers3=s.layers == 3) >= 0
                tiles = self.step3_tiles if ok else 0
            else:
                tiles = ops.net_step_pick_tiles(s.kind, s.F, s.h1, s.h2, d.max_n, d.max_k0, d.max_k1, d.max_e, s.Hd, s.out,
                                                layers3=s.layers == 3)
                # (more tiles than shared memory demands do not pay: measured on cfg3, B = 64, 1 / 2 / 4 tiles =
                # 41 / 46 / 82 us per step - the cluster barriers and DSMEM latency outweigh the halved rows)
            self._fused_fit[key] = tiles
        return tiles

    def _blob_only(self, d):
        """True when the step of batch ``d`` runs a cluster whole-step kernel, which stages the
        per-graph structure blobs and needs none of the global structure arrays: the structure pass
        is then the one-launch bitmap kernel (``ops.structure_blob``)."""
        s = self.spec
        if self._step3_tiles(d):
            key = (d.max_n, d.max_e, d.max_k0, d.max_k1, 'blobfit')
            fit = self._fused_fit.get(key)
            if fit is None:
                fit = ops.structure_blob_fits(d.max_n, d.max_e, d.max_k0, d.max_k1, weights=s.kind == 'sgat')
                self._fused_fit[key] = fit
            return bool(fit and self.blob_structure and not self.keep_intermediates)
        if not (self.blob_structure and self.fused_head and self.step_variant != 1 and not self.keep_intermediates
                and s.nb == 2 and d.cluster1 is not None and d.c1_ptr is not None and self._use_fused_graph(d)):
            return False
        key = (d.max_n, d.max_e, d.max_k0, d.max_k1, 'blob')
        fit = self._fused_fit.get(key)
        if fit is None:
            fit = (ops.ginet_step2_smem_bytes(s.F, s.h1, s.h2, d.max_n, d.max_k0, d.max_k1, d.max_e, s.Hd, s.out) >= 0
                   and ops.structure_blob_fits(d.max_n, d.max_e, d.max_k0, d.max_k1, weights=False)
                   and s.F % 4 == 0 and s.h1 % 4 == 0 and s.h2 % 4 == 0)
            self._fused_fit[key] = fit
        return fit

    def _use_fused_graph(self, d):
        s = self.spec
        if not (self.fused_graph and s.kind == 'ginet' and d.max_k0 and d.max_k1 and d.x.size(1) == s.F):
            return False
        key = (d.max_n, d.max_k0, d.max_k1)
        fit = self._fused_fit.get(key)
        if fit is None:
            fit = ops.ginet_fused_fits(s.F, s.h1, s.h2, s.nb, d.max_n, d.max_k0, d.max_k1)
            self._fused_fit[key] = fit
        return fit

    def _conv_aggregate(self, level, src, rowptr, col, Zin, n_rows, n_rows_dev, ew, s_out, post_out, tiles):
        s = self.spec
        cin = s.F if level == 0 else s.h1
        tile = {}
        if tiles is not None and src.stride(0) == src.size(1) and src.size(1) % 4 == 0:
            tile = dict(tile_ptr=tiles[0], tile_eptr=tiles[1], max_tile_rows=tiles[2], max_tile_edges=tiles[3])
        if s.kind == 'ginet':
            ops.aggregate(src, rowptr, col, Zin, n_rows=n_rows, n_rows_dev=n_rows_dev, **tile)
        elif s.kind == 'sgat':
            ops.aggregate(src, rowptr, col, Zin[:, cin:], C_=cin, ew=ew, self_src=src, self_out=Zin[:, :cin],
                          selfc_out=s_out, post_out=post_out, n_rows=n_rows, n_rows_dev=n_rows_dev, post_mode=1,
                          self_mode=2, **tile)
        else:
            ops.aggregate(src, rowptr, col, Zin[:, cin:], C_=cin, self_src=src, self_out=Zin[:, :cin],
                          post_out=post_out, n_rows=n_rows, n_rows_dev=n_rows_dev, post_mode=2, self_mode=1, **tile)

    def _forward(self, d, keep_mask=None, loss_inv=None):
        """``loss_inv``: when given (training step) the fused head also computes the loss and the
        gradients of fc1 / fc2 / the read-out (``_head_done``)."""
        s, ws, st, P = self.spec, self._ensure(d.B, d.N, d.E), self.structs[d.sslot], self.params
        N, B = d.N, d.B
        K0d, K1d = st.K0_dev, st.K1_dev
        pv = lambda name: P.view(P.data, name)
        flat = lambda name, n: P.data[P.offset(name):P.offset(name) + n]
        self._graph_done = self._all_done = self._adam_done = False
        tiles3 = self._step3_tiles(d)
        if tiles3:
            return self._forward_step3(d, st, tiles3, keep_mask, loss_inv)
        self._last_path = 'ops'
        if s.layers == 3:
            raise DrgnnError('the three-layer variant runs through the fused cluster kernel only; this batch does not fit it '
                             '(graphs of up to %d nodes / %d edges)' % (d.max_n, d.max_e))
        if self._use_fused_graph(d):
            # ONE launch: conv1 -> pool -> conv2 -> pool -> read-out, one CTA per graph (csrc/fused.cu)
            self._fa = ops.ginet_fused_args(st, d.x, flat('conv1.fc.weight', s.C1 * s.F),
                                            flat('conv2.fc.weight', s.nb * s.h2 * s.h1), ws.Zin1[:N], ws.Z1[:N],
                                            ws.arg0, ws.Zin2, ws.Z2, ws.arg1, ws.R[:B], d.node_ptr, B, s.F, s.h1, s.h2,
                                            s.nb, d.max_n, d.max_k0, d.max_k1, dR=ws.dR[:B], partial=ws.partial,
                                            dW1=self.grads[P.offset('conv1.fc.weight'):P.offset('conv1.fc.weight') + s.C1 * s.F],
                                            dW2=self.grads[P.offset('conv2.fc.weight'):P.offset('conv2.fc.weight') + s.nb * s.h2 * s.h1])
            key = (d.max_n, d.max_k0, d.max_k1, 'step')
            whole = self._fused_fit.get(key)
            if whole is None:
                whole = self.fused_head and ops.ginet_step_fits(s.F, s.h1, s.h2, s.nb, d.max_n, d.max_k0, d.max_k1,
                                                                  s.Hd, s.out)
                self._fused_fit[key] = whole
            if whole:
                # the whole step of every graph in ONE launch: forward, head, loss, backward (+ one reduction)
                drop = self.training and s.dropout > 0
                if drop and keep_mask is not None:
                    ws.keep[:B].copy_(keep_mask.to(self.device, F32))
                hashed = drop and keep_mask is None          # mask generated inside the kernel (counter-based hash)
                train_step = loss_inv is not None
                fuse_adam = train_step and self.world == 1 and self.fuse_adam
                use_comm = train_step and self.comm is not None and self._want_adam
                # several GPUs: the exchange over peer memory runs INSIDE the step kernel when every rank
                # launches the same co-resident grid (equal shards); else in its own launch behind it
                in_kernel = use_comm and not self._no_exchange and self._comm_in_kernel(d, st)
                task = ops.TASK_NONE
                if train_step:
                    task = ops.TASK_CE if self.task == 'class' else \
                        (ops.TASK_MSE_SIGMOID if self.transform_sigmoid else ops.TASK_MSE)
                    if self.task == 'class' and d.y_class is None:
                        raise DrgnnError('classification needs class-index targets (pass `classes` when building the batch)')
                    if self.task == 'reg' and d.y is None:
                        raise DrgnnError('the batch has no target')
                offs = [P.offset(nm) for nm in ('conv1.fc.weight', 'conv2.fc.weight', 'fc1.weight', 'fc1.bias',
                                                'fc2.weight', 'fc2.bias')]
                ops.ginet_step(self._fa, pv('fc1.weight'), pv('fc1.bias'), pv('fc2.weight'), pv('fc2.bias'), ws.pred[:B],
                               task=task, inv_norm=loss_inv if train_step else 1.0,
                               y=d.y if self.task == 'reg' else None, y_class=d.y_class if self.task == 'class' else None,
                               class_w=self.class_weights, keep=ws.keep[:B] if (drop and not hashed) else None,
                               keep_scale=1.0 / (1.0 - s.dropout) if drop else 1.0, loss=ws.loss,
                               partial=ws.partial_full, grads=self.grads, n_params=P.numel, offsets=offs,
                               forward_only=not train_step, drop_p=s.dropout if hashed else 0.0, seed=self.seed,
                               step_dev=self.step_dev,
                               adam=dict(p=P.data, m=self.exp_avg, v=self.exp_avg_sq, lr=self.lr, beta1=self.betas[0],
                                         beta2=self.betas[1], eps=self.eps) if (fuse_adam or in_kernel) else None,
                               skip_reduce=use_comm and not in_kernel, comm=self.comm if in_kernel else None,
                               max_e=d.max_e, mirror=self.keep_intermediates,
                               variant=2 if st.blob_only else self.step_variant, fuse_reduce=self.fuse_reduce,
                               blob=st.blob, gdesc=st.gstat if st.blob_only else None,
                               edge_ptr=d.edge_ptr, tc=self.fused_tc, timers=self.phase_timers,
                               zin1=getattr(st, 'zin1', None) if st.blob_only else None, head2=self.head_v2)
                self._graph_done = self._head_done = self._all_done = train_step
                self._all_done_kernel = True
                self._adam_done = fuse_adam
                if use_comm and not self._no_exchange:
                    self._last_exchange = 'in-kernel' if in_kernel else 'launch'
                if in_kernel:
                    self._reduced = self._adam_done = True
                elif use_comm:
                    # per-graph rows -> rank-local sum -> peers -> rank-ordered sum -> Adam: ONE launch
                    self._peer_exchange(partial=ws.partial_full, B=B)
                return ws.pred[:B]
            ops.ginet_fused_fwd(self._fa)
            self._graph_done = True
            return self._heads(d, keep_mask, loss_inv)
        # per-graph TMA pipeline for batches that fill the machine (>= 2 graphs per SM); small batches use
        # the L1-resident row kernel, whose launch latency is lower (3.0 us vs 4.1 us at B = 64)
        tiled = self.tiled and (d.B >= 2 * self._sms or self.tiled == 'force') and 8 * (d.max_n * s.F + d.max_n + 2 * d.max_e + 32) <= 200 * 1024
        # conv1: aggregate on the level-0 graph, then transform (+bias, ReLU)
        self._conv_aggregate(0, d.x, st.rowptr0, st.col0, ws.Zin1[:N], N, None, st.w0csr, ws.s0, ws.post0,
                             (d.node_ptr, d.edge_ptr, d.max_n, d.max_e) if tiled else None)
        if s.kind == 'ginet':
            W1, b1 = flat('conv1.fc.weight', s.C1 * s.F), None
            W2, b2 = flat('conv2.fc.weight', s.nb * s.h2 * s.h1), None
        elif s.kind == 'sgat':
            W1, b1 = flat('conv1.weight', 2 * s.F * s.h1), pv('conv1.bias')
            W2, b2 = flat('conv2.weight', 2 * s.h1 * s.h2), pv('conv2.bias')
        else:
            W1, b1 = flat('conv1.Wc', 2 * s.F * s.h1), pv('conv1.bias')
            W2, b2 = flat('conv2.Wc', 2 * s.h1 * s.h2), pv('conv2.bias')
        self._W1, self._W2 = W1, W2
        ops.linear(ws.Zin1[:N], W1, s.Kin1, s.C1, ws.Z1[:N], bias=b1, w_layout=s.w_layout, relu=True)
        # level-0 cluster max-pool (community_pooling.py:201)
        ops.maxpool_fwd(ws.Z1[:N], st.cmptr0, st.cmem0, ws.P1[:d.L1b], ws.arg0[:d.L1b], n_clusters_dev=K0d)
        # conv2 on the coarsened graph
        ew1 = st.edge_attr1.view(-1) if s.kind == 'sgat' else None
        L1 = d.L1b
        self._conv_aggregate(1, ws.P1[:L1], st.rowptr1, st.col1, ws.Zin2[:L1], L1, K0d, ew1, ws.s1, ws.post1, None)
        g2 = s.nb if s.kind == 'ginet' else 1
        ops.linear(ws.Zin2[:L1], W2, s.Kin2 // g2, s.h2, ws.Z2[:L1], bias=b2, groups=g2, w_layout=s.w_layout, relu=True,
                   rows_dev=K0d)
        # level-1 max-pool (max_pool_x) and graph read-out (scatter_mean by batch)
        ops.maxpool_fwd(ws.Z2[:L1], st.cmptr1, st.cmem1, ws.P2[:L1], ws.arg1[:L1], n_clusters_dev=K1d)
        ops.segment_mean_fwd(ws.P2[:L1], st.kptr1[:B + 1], ws.R[:B])
        return self._heads(d, keep_mask, loss_inv)

    def _forward_step3(self, d, st, tiles, keep_mask, loss_inv):
        """The whole step of every graph in ONE launch of the general cluster kernel (``ops.net_step``)."""
        s, ws, P, B = self.spec, self.ws, self.params, d.B
        drop = self.training and s.dropout > 0
        if drop and keep_mask is not None:
            ws.keep[:B].copy_(keep_mask.to(self.device, F32))
        hashed = drop and keep_mask is None
        train_step = loss_inv is not None
        fuse_adam = train_step and self.world == 1 and self.fuse_adam and self._want_adam
        use_comm = train_step and self.comm is not None and self._want_adam
        task = ops.TASK_NONE
        if train_step:
            task = ops.TASK_CE if self.task == 'class' else (ops.TASK_MSE_SIGMOID if self.transform_sigmoid else ops.TASK_MSE)
            if self.task == 'class' and d.y_class is None:
                raise DrgnnError('classification needs class-index targets (pass `classes` when building the batch)')
            if self.task == 'reg' and d.y is None:
                raise DrgnnError('the batch has no target')
        if s.kind == 'ginet':
            offs = dict(w1=P.offset('conv1.fc.weight'), w2=P.offset('conv2.fc.weight'))
        elif s.kind == 'sgat':
            offs = dict(w1=P.offset('conv1.weight'), b1=P.offset('conv1.bias'), w2=P.offset('conv2.weight'),
                        b2=P.offset('conv2.bias'))
        else:
            offs = dict(w1=P.offset('conv1.Wc'), b1=P.offset('conv1.bias'), w2=P.offset('conv2.Wc'), b2=P.offset('conv2.bias'))
        offs.update(fc1w=P.offset('fc1.weight'), fc1b=P.offset('fc1.bias'), fc2w=P.offset('fc2.weight'),
                    fc2b=P.offset('fc2.bias'))
        if s.layers == 3:
            offs.update(w3=P.offset('conv3.weight' if s.kind == 'sgat' else 'conv3.Wc'), b3=P.offset('conv3.bias'))
        in_kernel = False
        if use_comm and not self._no_exchange and self.fuse_comm and self.fuse_reduce and self._cur_B_global is not None \
                and d.B * self.world == self._cur_B_global:
            key = (d.max_n, d.max_e, d.max_k0, d.max_k1, tiles, 'clusters3')
            mc = self._fused_fit.get(key)
            if mc is None:
                smem = ops.net_step_smem_bytes(s.kind, tiles, s.F, s.h1, s.h2, d.max_n, d.max_k0, d.max_k1, d.max_e, s.Hd, s.out,
                                               layers3=s.layers == 3)
                mc = ops.net_step_max_clusters(s.kind, tiles, smem) if smem >= 0 else 0
                self._fused_fit[key] = mc
            in_kernel = d.B <= mc and tiles * s.nb * d.B <= int(self.comm.struct.max_blocks)
        mirror = None
        if self.keep_intermediates:
            mirror = dict(Zin1=ws.Zin1, Z1=ws.Z1, arg0=ws.arg0, Zin2=ws.Zin2, Z2=ws.Z2, arg1=ws.arg1)
        ops.net_step(s.kind, st, d.x, P.data, offs, B, s.F, s.h1, s.h2, s.Hd, s.out, d.max_n, d.max_e, d.max_k0, d.max_k1,
                     ws.pred[:B], d.node_ptr, d.edge_ptr, tiles=tiles, task=task,
                     inv_norm=loss_inv if train_step else 1.0, y=d.y if self.task == 'reg' else None,
                     y_class=d.y_class if self.task == 'class' else None, class_w=self.class_weights,
                     keep=ws.keep[:B] if (drop and not hashed) else None,
                     keep_scale=1.0 / (1.0 - s.dropout) if drop else 1.0, drop_p=s.dropout if hashed else 0.0,
                     seed=self.seed, loss=ws.loss, R=ws.R[:B], partial=ws.partial_full, grads=self.grads,
                     n_params=P.numel, forward_only=not train_step, step_dev=self.step_dev,
                     adam=dict(p=P.data, m=self.exp_avg, v=self.exp_avg_sq, lr=self.lr, beta1=self.betas[0],
                               beta2=self.betas[1], eps=self.eps) if (fuse_adam or in_kernel) else None,
                     skip_reduce=use_comm and not in_kernel, fuse_reduce=self.fuse_reduce,
                     comm=self.comm if in_kernel else None, mirror=mirror, tc=self.fused_tc,
                     timers=self.phase_timers, lpt=self.step3_lpt and tiles * s.nb * B > self._sm_count,
                     zin1=getattr(st, 'zin1', None) if st.blob_only else None)
        self._last_path = 'step3'
        self._graph_done = self._head_done = self._all_done = train_step
        self._all_done_kernel = True
        self._adam_done = fuse_adam
        if use_comm and not self._no_exchange:
            self._last_exchange = 'in-kernel' if in_kernel else 'launch'
        if in_kernel:
            self._reduced = self._adam_done = True
        elif use_comm:
            self._peer_exchange(partial=ws.partial_full, B=B)
        return ws.pred[:B]

    def _heads(self, d, keep_mask, loss_inv):
        s, ws, P, B = self.spec, self.ws, self.params, d.B
        pv = lambda name: P.view(P.data, name)
        # heads
        drop = self.training and s.dropout > 0
        scale = 1.0 / (1.0 - s.dropout) if drop else 1.0
        if drop:
            if keep_mask is not None:
                ws.keep[:B].copy_(keep_mask.to(self.device, F32))
            else:
                ws.keep[:B].bernoulli_(1.0 - s.dropout)
        self._head_done = False
        if self.fused_head and B <= 256 and self._head_fits:
            # ONE launch: fc1 -> fc2 -> loss -> backward of both (csrc/head.cu)
            task, grads = ops.TASK_NONE, {}
            if loss_inv is not None:
                task = ops.TASK_CE if self.task == 'class' else \
                    (ops.TASK_MSE_SIGMOID if self.transform_sigmoid else ops.TASK_MSE)
                if self.task == 'class' and d.y_class is None:
                    raise DrgnnError('classification needs class-index targets (pass `classes` when building the batch)')
                if self.task == 'reg' and d.y is None:
                    raise DrgnnError('the batch has no target')
                gv = lambda name: P.view(self.grads, name)
                grads = dict(dW1=gv('fc1.weight'), db1=gv('fc1.bias'), dW2=gv('fc2.weight'), db2=gv('fc2.bias'),
                             dR=ws.dR[:B])
            ops.head(ws.R[:B], pv('fc1.weight'), pv('fc1.bias'), pv('fc2.weight'), pv('fc2.bias'), ws.pred[:B],
                     task=task, inv_norm=loss_inv if loss_inv is not None else 1.0,
                     y=d.y if self.task == 'reg' else None, y_class=d.y_class if self.task == 'class' else None,
                     class_w=self.class_weights, keep=ws.keep[:B] if drop else None, keep_scale=scale,
                     loss=ws.loss if loss_inv is not None else None, **grads)
            self._head_done = loss_inv is not None
            return ws.pred[:B]
        ops.linear(ws.R[:B], pv('fc1.weight'), s.C2, s.Hd, ws.H[:B], bias=pv('fc1.bias'), relu=True,
                   out_mask=ws.keep[:B] if drop else None, mask_scale=scale)
        ops.linear(ws.H[:B], pv('fc2.weight'), s.Hd, s.out, ws.pred[:B], bias=pv('fc2.bias'))
        return ws.pred[:B]

    # ---------------------------------------------------------------- backward
    def _backward(self, d):
        if self._all_done:          # the whole-step kernel already produced every gradient
            return
        s, ws, st, P = self.spec, self.ws, self.structs[d.sslot], self.params
        N, B, L1 = d.N, d.B, d.L1b
        K0d = st.K0_dev
        pv = lambda name: P.view(P.data, name)
        gv = lambda name: P.view(self.grads, name)
        gflat = lambda name, n: self.grads[P.offset(name):P.offset(name) + n]
        drop = self.training and s.dropout > 0
        scale = 1.0 / (1.0 - s.dropout) if drop else 1.0
        # heads (already done by the fused head kernel in the forward when it applies)
        if not self._head_done:
            ops.linear_wgrad(ws.H[:B], ws.dpred[:B], s.Hd, s.out, gv('fc2.weight'), gv('fc2.bias'), work=ws.wwork)
            ops.linear(ws.dpred[:B], pv('fc2.weight'), s.out, s.Hd, ws.dH[:B], w_layout=1, out_mask=ws.H[:B],
                       mask_scale=scale)
            ops.linear_wgrad(ws.R[:B], ws.dH[:B], s.C2, s.Hd, gv('fc1.weight'), gv('fc1.bias'), work=ws.wwork)
            ops.linear(ws.dH[:B], pv('fc1.weight'), s.Hd, s.C2, ws.dR[:B], w_layout=1)
        if self._graph_done:
            # ONE launch (+ the partial sum): everything from dR down to dW1 / dW2, one CTA per graph
            ops.ginet_fused_bwd(self._fa)
            return
        # read-out and level-1 pool
        ops.segment_mean_bwd(ws.dR[:B], st.kptr1[:B + 1], ws.dP2[:L1])
        ops.maxpool_bwd(ws.dP2[:L1], ws.arg1[:L1], st.cl1, ws.dZ2[:L1], relu_out=ws.Z2[:L1], n_nodes_dev=K0d)
        # conv2
        g2 = s.nb if s.kind == 'ginet' else 1
        kin2 = s.Kin2 // g2
        if s.kind == 'ginet':
            dW2, db2 = gflat('conv2.fc.weight', s.nb * s.h2 * s.h1), None
            dW1, db1 = gflat('conv1.fc.weight', s.C1 * s.F), None
        elif s.kind == 'sgat':
            dW2, db2 = gflat('conv2.weight', 2 * s.h1 * s.h2), gv('conv2.bias')
            dW1, db1 = gflat('conv1.weight', 2 * s.F * s.h1), gv('conv1.bias')
        else:
            dW2, db2 = gflat('conv2.Wc', 2 * s.h1 * s.h2), gv('conv2.bias')
            dW1, db1 = gflat('conv1.Wc', 2 * s.F * s.h1), gv('conv1.bias')
        ops.linear_wgrad(ws.Zin2[:L1], ws.dZ2[:L1], kin2, s.h2, dW2, db2, groups=g2, w_layout=s.w_layout, rows_dev=K0d,
                         work=ws.wwork)
        ops.linear(ws.dZ2[:L1], self._W2, s.h2, kin2, ws.dZin2[:L1], groups=g2, w_layout=1 - s.w_layout, rows_dev=K0d)
        # transposed aggregation on the coarsened graph -> gradient of the pooled features
        if s.kind == 'ginet':
            ops.aggregate(ws.dZin2[:L1], st.cscptr1, st.cscrow1, ws.dP1[:L1], n_rows_dev=K0d)
        elif s.kind == 'sgat':
            ops.aggregate(ws.dZin2[:L1, s.h1:], st.cscptr1, st.cscrow1, ws.dP1[:L1], C_=s.h1, ew=st.w1csc,
                          sscale=ws.post1, self_src=ws.dZin2[:L1, :s.h1], selfc_in=ws.s1, self_mode=3, n_rows_dev=K0d)
        else:
            ops.aggregate(ws.dZin2[:L1, s.h1:], st.cscptr1, st.cscrow1, ws.dP1[:L1], C_=s.h1, sscale=ws.post1,
                          self_src=ws.dZin2[:L1, :s.h1], self_mode=1, n_rows_dev=K0d)
        # level-0 pool, conv1 (its input is data: only the weight gradient is needed)
        ops.maxpool_bwd(ws.dP1[:L1], ws.arg0[:L1], st.cl0, ws.dZ1[:N], relu_out=ws.Z1[:N])
        ops.linear_wgrad(ws.Zin1[:N], ws.dZ1[:N], s.Kin1, s.C1, dW1, db1, w_layout=s.w_layout, work=ws.wwork)

    def _adam(self):
        if self._adam_done:         # already applied by the fused reduction launch of the whole-step kernel
            self._adam_done = False
            return
        ops.adam_flat(self.params.data, self.grads, self.exp_avg, self.exp_avg_sq, self.step_dev, self.lr,
                      self.betas[0], self.betas[1], self.eps)

    # ---------------------------------------------------------------- public API
    def _inv_norm(self, d, B_global, inv_norm):
        if inv_norm is not None:
            return float(inv_norm)
        if self.task == 'reg':
            return 1.0 / float(d.B if B_global is None else B_global)
        raise DrgnnError("classification needs inv_norm = 1 / sum_b class_weight[target_b] over the GLOBAL batch "
                         "(CrossEntropyLoss(reduction='mean') normaliser, NeuralNet.py:258-263)")

    def _loss(self, d, inv_norm, with_grad=True):
        ws, B = self.ws, d.B
        if d.y is None and d.y_class is None:
            raise DrgnnError('the batch has no target')
        if self.task == 'reg':
            ops.mse_loss(ws.pred[:B].view(-1), d.y, inv_norm, ws.loss, ws.dpred[:B].view(-1) if with_grad else None,
                         sigmoid=self.transform_sigmoid)
        else:
            if d.y_class is None:
                raise DrgnnError('classification needs class-index targets (pass `classes` when building the batch)')
            ops.ce_loss(ws.pred[:B], d.y_class, inv_norm, ws.loss, ws.dpred[:B] if with_grad else None,
                        class_w=self.class_weights)
        return ws.loss

    def _peer_exchange(self, partial=None, B=0):
        if not self._no_exchange:
            ops.peer_reduce_adam(self.comm, self._grads_full, self.params.numel, self.params.numel + 4, partial=partial,
                                 B=B, step_dev=self.step_dev,
                                 adam=dict(p=self.params.data, m=self.exp_avg, v=self.exp_avg_sq, lr=self.lr,
                                           beta1=self.betas[0], beta2=self.betas[1], eps=self.eps))
        self._reduced = self._adam_done = True

    def _all_reduce(self):
        if self._reduced:               # the whole-step path already exchanged
            self._reduced = False
            return
        if self.world > 1 and self.comm is not None and self._want_adam:
            self._peer_exchange()
            self._reduced = False
            return
        if self.world > 1:
            # the path's only collective: one sum over ranks of [flat gradients | loss] (NCCL over NVLink)
            # (ws.loss IS the slot behind the gradients, see _ensure)
            if self.nccl is not None:
                self.nccl.all_reduce_(self._grads_full[:self.params.numel + 4])
            else:
                torch.distributed.all_reduce(self._grads_full[:self.params.numel + 4], group=self.pg)

    def forward(self, d, keep_mask=None, prepared=False):
        """Forward only (``model(batch)``): returns the ``[B, out]`` prediction (a view of an
        engine buffer, overwritten by the next call)."""
        if not prepared:
            self.prepare(d)
        return self._forward(d, keep_mask)

    def loss_and_grads(self, d, B_global=None, inv_norm=None, keep_mask=None):
        """Forward + loss + backward, no optimiser (parity tests).  Gradients in ``named_grads()``."""
        inv = self._inv_norm(d, B_global, inv_norm)
        self.prepare(d)
        self._want_adam = False
        try:
            self._forward(d, keep_mask, loss_inv=inv)
            if not self._head_done:
                self._loss(d, inv)
            self._backward(d)
        finally:
            self._want_adam = True
        return self.ws.loss, self.ws.pred[:d.B]

    def step(self, d, B_global=None, inv_norm=None, keep_mask=None, prepared=False):
        """One training step on a ``DeviceBatch``: structure pass (unless ``prepared``), forward,
        loss, backward, [all-reduce], Adam (the body of ``NeuralNet._epoch``'s loop,
        NeuralNet.py:490-503).  Returns (loss, pred) device tensors that the next step overwrites.
        ``B_global`` = graphs in the global batch when it is sharded over ranks: the local loss
        is sum/B_global so the all-reduced (summed) gradient is the gradient of the global mean
        (SURVEY 8e)."""
        inv = self._inv_norm(d, B_global, inv_norm)
        self._cur_B_global = B_global
        if self.use_graph and d.key is not None and keep_mask is None:
            if not prepared:
                self.prepare_graph(d)
            return self._step_graph(d, inv)
        if not prepared:
            self.prepare(d)
        self._forward(d, keep_mask, loss_inv=inv)
        if not self._head_done:
            self._loss(d, inv)
        self._backward(d)
        self._all_reduce()
        self._adam()
        return self.ws.loss, self.ws.pred[:d.B]

    # ---------------------------------------------------------------- packed batches / CUDA graphs
    def upload(self, pb, slot=0, sslot=None):
        """ONE host->device copy of a ``PackedBatch`` into a persistent device staging buffer
        (one per layout and slot, so CUDA-graph replays see fixed addresses).  Returns a
        DeviceBatch of views into it; its structure slot is ``sslot`` (default ``slot % STRUCT_SLOTS``)."""
        dev = self._staging.get((pb.layout_key(), slot))
        if dev is None:
            dev = torch.empty(pb.capacity_numel, dtype=F32, device=self.device)
            self._staging[(pb.layout_key(), slot)] = dev
        dev[:pb.numel].copy_(pb.buf, non_blocking=True)
        sslot = (slot % self.STRUCT_SLOTS) if sslot is None else int(sslot)
        ck = (pb.layout_key(), slot, sslot, pb.has_y)
        d = self._dbatch.get(ck)
        if d is None:
            # the views depend on the layout only (cluster1 is a capacity-sized view, its live length is
            # d.L1), so one DeviceBatch per staging slot is built once and re-used: no per-step tensor views
            d = DeviceBatch.from_packed(pb, dev)
            d.sslot = sslot
            d.key = (pb.layout_key(), slot, sslot)
            self._dbatch[ck] = d
        d.L1 = pb.L1
        d.mol = pb.mol
        return d

    def _capture(self, fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            fn()
        return g

    def prepare_graph(self, d):
        """``prepare`` replayed from a CUDA graph (captured on first use for the batch's layout)."""
        key = ('prep', d.key)
        g = self._graphs.get(key)
        self._primed = None
        if g is None:
            self.prepare(d)                     # eager warm-up (first-use attribute setup, buffer growth)
            if self._graphs.get(key) is None:   # growth clears the cache; the warm-up result stays valid
                cur = torch.cuda.current_stream(self.device)
                g = self._capture(lambda: self.prepare(d))
                cur.synchronize()
                self._graphs[key] = g
            return
        g.replay()

    def _prep_graph_handle(self, d):
        """The captured structure-pass graph of batch ``d``'s slot (captured now if needed; the pass runs
        once eagerly on the way, which is idempotent)."""
        key = ('prep', d.key)
        if self._graphs.get(key) is None:
            self.prepare_graph(d)
            if self._graphs.get(key) is None:   # buffer growth dropped it: second try captures
                self.prepare_graph(d)
        return self._graphs[key]

    def _step_graph(self, d, inv):
        """Replay (capture on first use) everything after the structure pass as one CUDA graph.
        The graph is tied to the staging buffer / structure slot of the batch; with several ranks
        the gradient all-reduce sits between the two captured halves."""
        g1, g2 = self._step_graph_entry(d, inv)
        g1.replay()
        if g2 is not None:
            self._all_reduce()
            g2.replay()
        return self.ws.loss, self.ws.pred[:d.B]

    def _step_graph_entry(self, d, inv):
        """The captured step graph(s) of batch ``d``'s slot: ``(g1, g2)``, captured on first use (warm-up on
        the slot's current contents with the optimiser state restored afterwards); nothing is replayed."""
        key = ('step', d.key, round(inv, 12), self.training)
        ent = self._graphs.get(key)
        if ent is None:
            # warm up un-captured (first-use attribute setup), then capture
            snap = [t.clone() for t in (self.params.data, self.exp_avg, self.exp_avg_sq, self.step_dev)]
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            split = self.world > 1 and self.comm is None      # NCCL path: the all-reduce sits between two graphs
            with torch.cuda.stream(side):
                self._no_exchange = True                      # ranks may warm up at different steps: stay local
                try:
                    self._forward(d, loss_inv=inv)
                    if not self._head_done:
                        self._loss(d, inv)
                    self._backward(d)
                    if not split:
                        self._all_reduce()
                    self._adam()
                finally:
                    self._no_exchange = False
            torch.cuda.current_stream(self.device).wait_stream(side)
            for t, c in zip((self.params.data, self.exp_avg, self.exp_avg_sq, self.step_dev), snap):
                t.copy_(c)

            def body():
                self._forward(d, loss_inv=inv)
                if not self._head_done:
                    self._loss(d, inv)
                self._backward(d)
                if not split:
                    self._all_reduce()
                    self._adam()
            g1 = self._capture(body)
            g2 = self._capture(self._adam) if split else None
            ent = (g1, g2)
            self._graphs[key] = ent
        return ent

    def _eval_graph_entry(self, d, inv):
        """The captured scoring graph of batch ``d``'s slot: forward (+ loss without gradient), captured on
        first use; nothing is replayed.  The inner call of ``NeuralNet.eval`` / ``test`` (NeuralNet.py:432-460)."""
        key = ('eval', d.key, round(inv, 12), self.training)
        g = self._graphs.get(key)
        if g is None:
            def body():
                self._forward(d)
                self._loss(d, inv, with_grad=False)
            side = torch.cuda.Stream(self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):
                body()                                   # eager warm-up: first-use setup inside the C-ABI
            torch.cuda.current_stream(self.device).wait_stream(side)
            g = self._capture(body)
            self._graphs[key] = g
        return g

    def _pipeline_state(self):
        if self._copy_stream is None:
            ns = self.STRUCT_SLOTS
            self._copy_stream = torch.cuda.Stream(self.device)
            # two structure streams: the passes of batches i+1 and i+2 run side by side (they are latency
            # bound, one CTA per graph, and leave most SMs idle)
            self._prep_streams = [torch.cuda.Stream(self.device), torch.cuda.Stream(self.device)]
            self._prep_stream = self._prep_streams[0]
            ev = torch.cuda.Event
            self._slot_free = [ev() for _ in range(ns)]     # structure slot released by the step that read it
            self._slot_ready = [ev() for _ in range(ns)]    # structure pass of the slot finished
            self._stage_free = [ev() for _ in range(ns)]    # staging buffer released by the step that read it
            self._stage_copied = [ev() for _ in range(ns)]  # H2D copy into the staging buffer finished
        return self._copy_stream

    def _prepare_any(self, d):
        if self.use_graph and d.key is not None:
            self.prepare_graph(d)
        else:
            self.prepare(d)

    def train_batches(self, packed_batches, B_global=None, inv_norms=None, train=True):
        """Pipelined pass over ``PackedBatch`` objects held in (pinned) host memory - the inner
        loop of ``NeuralNet._epoch`` / ``eval`` (NeuralNet.py:490-523, 432-460).  Per batch, on a
        side stream: ONE host->device copy and the structure pass (two staging / structure slots,
        so both overlap the compute of the previous batch); on the main stream: the fused step
        (or forward + loss when ``train=False``) and an asynchronous device->host read of
        ``[loss, pred...]`` into pinned memory.  One host synchronisation at the end.
        Returns (losses [n], preds list)."""
        main = torch.cuda.current_stream(self.device)
        cs = self._pipeline_state()
        ns = self.STRUCT_SLOTS
        self._primed = None
        cs.wait_stream(main)
        for ps in self._prep_streams:
            ps.wait_stream(main)
        outs = []
        was_training = self.training
        self.train(train)
        packed_batches = list(packed_batches)
        # one pinned read-back block for the whole pass (a pinned allocation per step would cost more than the step)
        width = 4 + max([pb.B for pb in packed_batches] + [1]) * self.spec.out
        rows = max(len(packed_batches), 1)
        # the pinned block is cached (grow-only): cudaHostAlloc costs 0.1-1 ms and serialises between the
        # processes of a multi-GPU job - inside a short pass it was the N = 4 / 8 end-to-end outlier
        if self._host_out is None or self._host_out.numel() < rows * width:
            self._host_out = torch.empty(max(rows * width, 4096), dtype=F32, pin_memory=True)
        host_all = self._host_out[:rows * width].view(rows, width)
        if self._feed_native(packed_batches, B_global, inv_norms, train, host_all, main, cs):
            main.synchronize()
            self.train(was_training)
            n0, B0, out = self.params.numel, packed_batches[0].B, self.spec.out
            losses = host_all[:, 0].clone()
            pall = host_all[:, 4:4 + B0 * out].clone()          # the pinned block is reused by the next pass
            preds = [pall[i].view(B0, out) for i in range(len(packed_batches))]
            return losses, preds
        for i, pb in enumerate(packed_batches):
            # pipeline: H2D copies and structure passes of the next batches (up to STRUCT_SLOTS - 1 ahead, on the
            # copy stream and two alternating structure streams) | step of batch i
            stg = slot = i % ns
            ps = self._prep_streams[i & 1]
            with torch.cuda.stream(cs):
                if i >= ns:
                    cs.wait_event(self._stage_free[stg])
                d = self.upload(pb, stg, slot)
                self._stage_copied[stg].record(cs)
            with torch.cuda.stream(ps):
                ps.wait_event(self._stage_copied[stg])
                if i >= ns:
                    ps.wait_event(self._slot_free[slot])
                self._prepare_any(d)
                self._slot_ready[slot].record(ps)
            main.wait_event(self._slot_ready[slot])
            inv = None if inv_norms is None else inv_norms[i]
            if train:
                loss, pred = self.step(d, B_global=B_global, inv_norm=inv, prepared=True)
            else:
                pred = self._forward(d)
                loss = self._loss(d, self._inv_norm(d, B_global, inv), with_grad=False) \
                    if (d.y is not None or d.y_class is not None) else self.ws.loss
            # [loss, pad(3), pred...] is contiguous behind the flat gradients: ONE device->host copy
            n0 = self.params.numel
            host = host_all[i, :4 + pred.numel()]
            host.copy_(self._grads_full[n0:n0 + 4 + pred.numel()], non_blocking=True)
            self._slot_free[slot].record(main)
            self._stage_free[stg].record(main)
            outs.append((host, tuple(pred.shape)))
        main.synchronize()
        self.train(was_training)
        losses = torch.stack([h[0] for h, _ in outs]) if outs else torch.zeros(0)
        preds = [h[4:].clone().view(shape) for h, shape in outs]
        return losses, preds

    def _feed_native(self, packed_batches, B_global, inv_norms, train, host_all, main, cs):
        """The loop of ``train_batches`` issued from C (``drgnn_feed_run``, csrc/feed.cu) when every batch
        shares one layout (fixed-shape mini-batches), trains through the cluster step kernel and uses one
        loss normaliser: per step the host then pays a handful of runtime calls instead of ~20 calls into
        torch (55 us, more than the step and the PCIe copy need).  Returns False when not applicable."""
        import ctypes as C
        from . import _lib
        ns = self.STRUCT_SLOTS
        n = len(packed_batches)
        if not (self.native_feed and self.use_graph and n > ns and (self.world == 1 or self.comm is not None or not train)):
            return False
        key0 = packed_batches[0].layout_key()
        if any(pb.layout_key() != key0 or not pb.has_y for pb in packed_batches):
            return False
        if inv_norms is not None and any(abs(v - inv_norms[0]) > 0 for v in inv_norms):
            return False
        if not hasattr(torch.cuda.CUDAGraph, 'raw_cuda_graph_exec'):
            return False
        inv0 = None if inv_norms is None else inv_norms[0]
        # one staging slot, structure slot and pair of captured graphs per pipeline slot (cached per layout)
        ck = ('feed', key0, B_global, inv0, self.training, bool(train))
        slots = self._graphs.get(ck)
        if slots is None:
            slots = []
            for j in range(ns):
                d = self.upload(packed_batches[j], j, j)
                if not self._blob_only(d):
                    return False
                inv = self._inv_norm(d, B_global, inv0)
                self._cur_B_global = B_global
                pg = self._prep_graph_handle(d)
                g1, g2 = self._step_graph_entry(d, inv) if train else (self._eval_graph_entry(d, inv), None)
                if g2 is not None:
                    return False
                slots.append((d, self._staging[(key0, j)], pg, g1, pg.raw_cuda_graph_exec(), g1.raw_cuda_graph_exec()))
            torch.cuda.current_stream(self.device).synchronize()
            if any(self._graphs.get(('prep', sl[0].key)) is not sl[2] for sl in slots):
                return False             # buffer growth dropped a graph meanwhile: take the Python loop this time
            self._graphs[ck] = slots
        import numpy as np
        n0 = self.params.numel
        out_bytes = 4 * (4 + packed_batches[0].B * self.spec.out)
        d_out = self._grads_full.data_ptr() + 4 * n0
        # drgnn_feed_step records filled column-wise (9 x int64 per step; the last word holds slot | reserved)
        rec = np.empty((n, 9), dtype=np.int64)
        sl_idx = np.arange(n, dtype=np.int64) % ns
        rec[:, 0] = [pb.buf.data_ptr() for pb in packed_batches]
        rec[:, 1] = np.asarray([sl[1].data_ptr() for sl in slots], dtype=np.int64)[sl_idx]
        rec[:, 2] = [4 * pb.numel for pb in packed_batches]
        rec[:, 3] = np.asarray([sl[4] for sl in slots], dtype=np.int64)[sl_idx]
        rec[:, 4] = np.asarray([sl[5] for sl in slots], dtype=np.int64)[sl_idx]
        rec[:, 5] = d_out
        rec[:, 6] = host_all.data_ptr() + np.arange(n, dtype=np.int64) * (host_all.stride(0) * 4)
        rec[:, 7] = out_bytes
        rec[:, 8] = sl_idx
        assert C.sizeof(_lib.FeedStep) == 72
        steps = rec.ctypes.data_as(C.POINTER(_lib.FeedStep))
        if self._read_stream is None:
            self._read_stream = torch.cuda.Stream(self.device)
        ring_slots, ring_stride = 8, (out_bytes + 255) // 256 * 256
        if self._read_ring is None or self._read_ring.numel() < ring_slots * ring_stride:
            self._read_ring = torch.zeros(ring_slots * ring_stride, dtype=torch.uint8, device=self.device)
        self._feed_keep = (rec, packed_batches, host_all)         # alive until the streams have drained
        t_issue = time.perf_counter()
        _lib.check(_lib.load().drgnn_feed_run(steps, n, ns, main.cuda_stream, cs.cuda_stream,
                                              self._prep_streams[0].cuda_stream, self._prep_streams[1].cuda_stream,
                                              self._read_stream.cuda_stream, self._read_ring.data_ptr(), ring_stride,
                                              ring_slots),
                   'drgnn_feed_run')
        self.feed_issue_us = 1e6 * (time.perf_counter() - t_issue) / n      # host cost of issuing one step
        last = slots[(n - 1) % ns][0]
        last.L1, last.mol = packed_batches[-1].L1, packed_batches[-1].mol
        self._last_struct = self.structs[last.sslot]
        return True

    def train_resident(self, dbatches, steps=None, B_global=None, start=0):
        """Training steps over batches already resident in HBM (``upload``-ed DeviceBatches, e.g. a
        data set cached on the device across epochs), cycling through ``dbatches`` for ``steps``
        steps.  The structure passes of the next batches run on two side streams while step i
        computes (``upload(pb, slot)`` spreads the batches over ``STRUCT_SLOTS`` structure slots;
        a slot is rewritten only after the step that read it).  No host synchronisation.  Returns
        (loss, pred) of the last step.  ``start``: index of the first batch of this call in the rotation;
        a call that continues where the previous one stopped finds its first structure passes done."""
        n = len(dbatches) if steps is None else steps
        start = int(start) % max(len(dbatches), 1)
        main = torch.cuda.current_stream(self.device)
        self._pipeline_state()
        out = None
        first = 0
        R = len(dbatches)
        C = self._rotation_chunk(dbatches) if (self.use_graph and self.rotation_graph and n >= 1 and
                                               (self.world == 1 or self.comm is not None)) else 0
        if C:
            # CUDA graphs of C consecutive steps each (structure passes of the batches two steps ahead on side
            # streams + the steps, with their dependencies): the host issues one launch per C steps instead of
            # ~4 calls per step - the per-step issue cost had become the bound of the resident loop.  The
            # remainder n % C is one more (shorter) chunk graph, so NO step is issued eagerly.
            LA = self.ROTATION_LOOKAHEAD
            rot = tuple(d.key for d in dbatches)
            # (a first-use capture runs an eager warm-up step that rewrites structure slots: capture first, prime after)
            graphs = [self._chunk_graph(dbatches, (start + c * C) % R, C, B_global) for c in range(n // C)]
            first = (n // C) * C
            if n > first:
                graphs.append(self._chunk_graph(dbatches, (start + first) % R, n - first, B_global))
            if self._primed != (rot, start):
                for j in range(LA):                  # structure passes the first chunk expects to find done
                    self._prepare_any(dbatches[(start + j) % R])
            for g in graphs:
                g.replay()
            self._primed = (rot, (start + n) % R)    # every chunk leaves the passes of the next LA batches done
            return (self.ws.loss, self.ws.pred[:dbatches[(start + n - 1) % R].B])
        for ps in self._prep_streams:
            ps.wait_stream(main)
        used = set()
        for i in range(first, n):
            d = dbatches[(start + i) % len(dbatches)]
            slot = d.sslot
            ps = self._prep_streams[i & 1]
            with torch.cuda.stream(ps):
                if slot in used:
                    ps.wait_event(self._slot_free[slot])       # the step that last read this slot has finished
                self._prepare_any(d)
                self._slot_ready[slot].record(ps)
            main.wait_event(self._slot_ready[slot])
            out = self.step(d, B_global=B_global, prepared=True)
            self._slot_free[slot].record(main)
            used.add(slot)
        return out

    ROTATION_LOOKAHEAD = 2      # structure passes run this many steps ahead inside a chunk graph (< STRUCT_SLOTS)

    def _rotation_chunk(self, dbatches):
        """Steps per chunk graph for ``train_resident`` (0: not applicable); see ``rotation_chunk``."""
        if any(d.key is None for d in dbatches):
            return 0
        return rotation_chunk([d.sslot for d in dbatches], self.STRUCT_SLOTS, self.ROTATION_LOOKAHEAD)

    def _chunk_graph(self, dbatches, start, C, B_global):
        """One CUDA graph: steps ``start .. start+C-1`` of the rotation on the capturing stream and the
        structure passes of batches ``start+LA .. start+C-1+LA`` on two side streams.  Pass j waits for step
        j - STRUCT_SLOTS (the last reader of its slot) when that step is in this chunk; step i waits for pass
        i when that pass is in this chunk - everything older finished with the previous replay."""
        key = ('chunk', tuple(d.key for d in dbatches), start, C, B_global, self.training)
        g = self._graphs.get(key)
        if g is not None:
            return g
        R, ns, la = len(dbatches), self.STRUCT_SLOTS, self.ROTATION_LOOKAHEAD
        for d in dbatches:
            self._ensure(d.B, d.N, d.E)
        # eager warm-up of one step (first-use setup inside the C-ABI), optimiser state restored afterwards
        snap = [t.clone() for t in (self.params.data, self.exp_avg, self.exp_avg_sq, self.step_dev)]
        use_graph, self.use_graph = self.use_graph, False
        try:
            self._no_exchange = True
            try:
                self.step(dbatches[start % R], B_global=B_global)
                for j in range(la):              # the warm-up rewrote a structure slot: restore what the chunk expects
                    self.prepare(dbatches[(start + j) % R])
            finally:
                self._no_exchange = False
            torch.cuda.current_stream(self.device).synchronize()
            for t, c in zip((self.params.data, self.exp_avg, self.exp_avg_sq, self.step_dev), snap):
                t.copy_(c)
            done, ready = {}, {}
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                main = torch.cuda.current_stream(self.device)
                for ps in self._prep_streams:
                    ps.wait_stream(main)
                for i in range(start, start + C):
                    j = i + la
                    ps = self._prep_streams[j & 1]
                    with torch.cuda.stream(ps):
                        if j - ns in done:
                            ps.wait_event(done[j - ns])
                        self.prepare(dbatches[j % R])
                        ready[j] = torch.cuda.Event()
                        ready[j].record(ps)
                    if i in ready:
                        main.wait_event(ready[i])
                    self.step(dbatches[i % R], B_global=B_global, prepared=True)
                    done[i] = torch.cuda.Event()
                    done[i].record(main)
                for ps in self._prep_streams:
                    main.wait_stream(ps)
        finally:
            self.use_graph = use_graph
        self._graphs[key] = g
        return g

    def validate(self):
        """Raise if ANY structure pass / step since the last call flagged invalid input (ONE host sync).
        Every structure slot keeps a sticky status word (never cleared by a launch), so a malformed batch in
        the middle of an epoch is reported by the check at its end.  After an error the grid-barrier counters
        of the fused step are re-armed, so training may go on with the next batch."""
        live = [st for st in self.structs if st is not None]
        if live:
            bits = 0
            for v in torch.cat([st.status for st in live]).cpu().tolist():
                bits |= int(v)
            for st in live:
                st._counts_host = None
            if bits:
                for st in live:
                    st.status.zero_()
                self.step_dev[1:3].zero_()       # ticket / barrier counter of the in-kernel reduction
                from . import _lib
                msgs = [t for bit, t in _lib.STATUS_TEXT.items() if bits & bit]
                raise DrgnnError('invalid batch structure: ' + '; '.join(msgs or ['status %d' % bits]))
        if self.comm is not None and self.comm.status() != 0:
            raise DrgnnError('gradient exchange: a peer rank did not deliver its gradients within the watchdog '
                             '(DRGNN_PEER_TIMEOUT_S); the weights of this rank are no longer valid')
