"""Data-parallel sharding of a mini-batch of independent graphs.

Graphs of a mini-batch are independent (block-diagonal ``edge_index``, per-graph clusters,
per-graph read-out), so the path shards with NO data-path collective: every rank runs the
fused step on its own graphs and the only exchange is ONE all-reduce (sum) of the flat
gradient buffer before Adam.  The loss of a rank is ``sum_local(...) / B_global`` so the
summed gradients equal the gradient of the global mean whatever the shard sizes.
"""
import os

import torch


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(','):
        if not part:
            continue
        a, _, b = part.partition('-')
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE any pinned host memory is
    allocated (first-touch: the feeder's pinned blocks then live in the memory next to the GPU's PCIe root
    instead of crossing the socket interconnect - with 8 feeders that link is the end-to-end bound).
    Reads /sys (numa_node of the GPU's PCI function, cpulist of the node); returns a short description of
    what was done, never raises."""
    try:
        props = torch.cuda.get_device_properties(device_index)
        bdf = None
        if all(hasattr(props, k) for k in ('pci_domain_id', 'pci_bus_id', 'pci_device_id')):
            bdf = '%04x:%02x:%02x.0' % (int(props.pci_domain_id), int(props.pci_bus_id), int(props.pci_device_id))
        if bdf is None:
            import ctypes
            rt = ctypes.CDLL('libcudart.so')
            buf = ctypes.create_string_buffer(32)
            if rt.cudaDeviceGetPCIBusId(buf, 32, int(device_index)) != 0:
                return 'unbound (no PCI bus id)'
            bdf = buf.value.decode()
        if isinstance(bdf, int):
            return 'unbound (PCI bus id unavailable)'
        bdf = bdf.lower()
        if len(bdf.split(':')[0]) == 8:           # cudart prints an 8-digit domain, sysfs uses 4
            bdf = bdf[4:]
        node_path = '/sys/bus/pci/devices/%s/numa_node' % bdf
        if not os.path.exists(node_path):
            return 'unbound (%s missing)' % node_path
        with open(node_path) as f:
            node = int(f.read().strip())
        if node < 0:
            return 'unbound (single NUMA node)'
        with open('/sys/devices/system/node/node%d/cpulist' % node) as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        cpus = (cpus & allowed) or allowed
        os.sched_setaffinity(0, cpus)
        return 'GPU %s on NUMA node %d: bound to %d CPUs' % (bdf, node, len(cpus))
    except Exception as e:                         # pragma: no cover - best effort
        return 'unbound (%s)' % e


def _gpu_numa_node(device_index):
    """NUMA node of the GPU's PCI function (from /sys), or None."""
    props = torch.cuda.get_device_properties(device_index)
    if not all(hasattr(props, k) for k in ('pci_domain_id', 'pci_bus_id', 'pci_device_id')):
        return None
    bdf = '%04x:%02x:%02x.0' % (int(props.pci_domain_id), int(props.pci_bus_id), int(props.pci_device_id))
    node_path = '/sys/bus/pci/devices/%s/numa_node' % bdf
    if not os.path.exists(node_path):
        return None
    with open(node_path) as f:
        node = int(f.read().strip())
    return node if node >= 0 else None


def prefer_gpu_numa_memory(device_index, enable=True):
    """Memory policy of the calling thread: prefer the NUMA node the GPU hangs off (``enable=False``: back to the
    default policy).  Pinned host blocks allocated afterwards (``PackedBatch``) then live next to the GPU's PCIe
    root - the same H2D copy is 15-40 % slower out of the other socket's memory - WITHOUT touching the CPU affinity
    (a single-process run keeps every core, e.g. for the CPU baseline timed in the same process).  Linux
    ``set_mempolicy`` through ctypes; returns a short description, never raises."""
    try:
        import ctypes
        libc = ctypes.CDLL(None, use_errno=True)
        SYS_set_mempolicy, MPOL_DEFAULT, MPOL_PREFERRED = 238, 0, 1          # x86-64
        import platform
        if platform.machine() not in ('x86_64', 'AMD64'):
            return 'memory policy unchanged (not x86-64)'
        if not enable:
            rc = libc.syscall(SYS_set_mempolicy, MPOL_DEFAULT, None, 0)
            return 'default memory policy' if rc == 0 else 'memory policy unchanged (errno %d)' % ctypes.get_errno()
        node = _gpu_numa_node(device_index)
        if node is None:
            return 'memory policy unchanged (single NUMA node)'
        mask = (ctypes.c_ulong * 16)()
        mask[node // 64] = 1 << (node % 64)
        rc = libc.syscall(SYS_set_mempolicy, MPOL_PREFERRED, mask, 16 * 64 + 1)
        if rc != 0:
            return 'memory policy unchanged (errno %d)' % ctypes.get_errno()
        return 'single process: host memory preferred on NUMA node %d (the GPU\'s), CPU affinity unchanged' % node
    except Exception as e:                         # pragma: no cover - best effort
        return 'memory policy unchanged (%s)' % e


def graph_cost(d):
    """Work estimate of one graph: nodes + directed edges."""
    return int(d.num_nodes) + int(d.num_edges)


def shard_indices(costs, world, balance=True):
    """Partition graph indices ``0..len(costs)-1`` into ``world`` shards.

    ``balance=False``: contiguous equal-count chunks (fixed-size configurations).
    ``balance=True``: longest-processing-time bin packing on ``costs`` with shard sizes kept
    within one graph of each other (mixed-size batches, BASELINE config 5).  Deterministic."""
    n = len(costs)
    if world <= 0:
        raise ValueError('world must be positive')
    if not balance:
        base, rem = divmod(n, world)
        out, s = [], 0
        for r in range(world):
            k = base + (1 if r < rem else 0)
            out.append(list(range(s, s + k)))
            s += k
        return out
    cap = (n + world - 1) // world
    order = sorted(range(n), key=lambda i: (-costs[i], i))
    loads = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min((r for r in range(world) if len(out[r]) < cap), key=lambda r: (loads[r], r))
        out[r].append(i)
        loads[r] += costs[i]
    return [sorted(s) for s in out]


def shard_graphs(graphs, world, rank, balance=True):
    parts = shard_indices([graph_cost(g) for g in graphs], world, balance)
    return [graphs[i] for i in parts[rank]]


def init_distributed(backend=None):
    """One process per GPU (torchrun env: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, world, local_rank).  NCCL on CUDA, gloo otherwise."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not torch.distributed.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            torch.distributed.init_process_group(backend, device_id=torch.device('cuda', local))
        else:
            torch.distributed.init_process_group(backend)
    return rank, world, local


def all_reduce_flat_(flat, group=None):
    """The path's only collective: sum the flat gradient buffer over ranks, in place."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(flat, group=group)
    return flat


class PeerComm(object):
    """Exchange region of the fused gradient all-reduce (``csrc/comm.cu``, ``include/drgnn.h``).

    Every rank allocates one region through the library (cudaMalloc + CUDA IPC handle), the
    handles are exchanged with ``all_gather_object`` and every rank maps the regions of its peers,
    after which ``ops.peer_reduce_adam`` moves gradients with plain stores over NVLink - no NCCL
    call on the step.  ``regions`` (same-process pointers) builds a communicator without IPC: used
    by the single-GPU protocol test, where two "ranks" run on two streams of one device.
    Region layout: ctr[16] u32 | flags [2][world][max_blocks] u32 | buffers [2][world][stride] f32 |
    low-latency slots [2][world][stride] x 8 bytes (value + epoch in one store; in-kernel exchange)."""

    THREADS = 256

    @staticmethod
    def layout(world, n_sum):
        stride = (int(n_sum) + 3) // 4 * 4
        # one flag per block of the stand-alone kernel (256 elements each) or per CTA of the cluster step
        # kernel when the exchange runs inside it (at most 2 CTAs per SM pair, 320 covers every B200)
        max_blocks = max((int(n_sum) + PeerComm.THREADS - 1) // PeerComm.THREADS, 320)
        flags_off = 64
        buf_off = (flags_off + 4 * 2 * world * max_blocks + 255) // 256 * 256
        ll_off = (buf_off + 4 * 2 * world * stride + 255) // 256 * 256     # 8-byte {value, epoch} slots
        return stride, max_blocks, flags_off, buf_off, ll_off, ll_off + 8 * 2 * world * stride

    def __init__(self, n_sum, rank=None, world=None, group=None, regions=None, timeout_s=None):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        dist = torch.distributed
        if regions is None:
            world = dist.get_world_size(group)
            rank = dist.get_rank(group)
        if world > _lib.MAX_PEERS:
            raise _lib.DrgnnError('PeerComm supports up to %d ranks (one NVSwitch domain)' % _lib.MAX_PEERS)
        self.world, self.rank, self.n_sum = int(world), int(rank), int(n_sum)
        stride, max_blocks, flags_off, buf_off, ll_off, nbytes = self.layout(world, n_sum)
        self.nbytes = nbytes
        self._own = None
        self._opened = []
        if regions is None:
            # every collective below is entered by ALL ranks whatever failed locally, and the outcome is
            # agreed on: either every rank ends up with a communicator or every rank raises
            own = _lib.VP()
            handle = C.create_string_buffer(_lib.IPC_HANDLE_BYTES)
            err = None
            if lib.drgnn_comm_alloc(nbytes, C.byref(own), handle) != 0:
                err = 'rank %d: %s' % (rank, lib.drgnn_last_error().decode())
            else:
                self._own = own.value
            handles = [None] * world
            dist.all_gather_object(handles, None if err else handle.raw, group=group)
            regions = []
            if all(h is not None for h in handles):
                for r in range(world):
                    if r == rank:
                        regions.append(self._own)
                        continue
                    p = _lib.VP()
                    if lib.drgnn_comm_open(handles[r], C.byref(p)) != 0:
                        err = 'rank %d -> %d: %s' % (rank, r, lib.drgnn_last_error().decode())
                        break
                    self._opened.append(p.value)
                    regions.append(p.value)
            else:
                err = err or 'a peer could not allocate its exchange region'
            errs = [None] * world
            dist.all_gather_object(errs, err, group=group)
            if any(e is not None for e in errs):
                self.close()
                raise _lib.DrgnnError('peer-memory exchange unavailable: ' + '; '.join(e for e in errs if e))
        self.regions = list(regions)
        st = _lib.PeerComm()
        st.world, st.rank = self.world, self.rank
        for r in range(self.world):
            st.xflag[r] = self.regions[r] + flags_off
            st.xbuf[r] = self.regions[r] + buf_off
            st.xll[r] = self.regions[r] + ll_off
        st.ctr = self.regions[self.rank]
        st.stride, st.max_blocks = stride, max_blocks
        if timeout_s is None:
            timeout_s = float(os.environ.get('DRGNN_PEER_TIMEOUT_S', '20'))
        st.timeout_ns = int(timeout_s * 1e9)
        self.struct = st

    def status(self):
        """ctr[2] of the own region: non-zero if a peer failed to deliver within the watchdog."""
        import ctypes as C
        from . import _lib
        out = (C.c_uint32 * 4)()
        _lib.check(_lib.load().drgnn_comm_status(self.regions[self.rank], out), 'drgnn_comm_status')
        return int(out[2])

    def close(self):
        from . import _lib
        lib = _lib.load()
        for p in self._opened:
            lib.drgnn_comm_close(p)
        self._opened = []
        if self._own is not None:
            lib.drgnn_comm_free(self._own)
            self._own = None


class NcclComm(object):
    """The path's one collective through the C-ABI (``drgnn_nccl_*``, ``csrc/nccl_bridge.cu``): an NCCL
    communicator owned by libdrgnn - the fallback of the gradient exchange where peer memory cannot be
    mapped, and the form a non-PyTorch host would bind.

    The 128-byte unique id of rank 0 reaches the other ranks through ``torch.distributed``
    (``broadcast_object_list`` over ``group``; any backend) or, without a process group, through
    ``id_bytes`` handed in by the launcher.  ``all_reduce_(t)`` sums a contiguous fp32 CUDA tensor
    over the ranks in place on the current stream."""

    def __init__(self, rank=None, world=None, group=None, id_bytes=None):
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        if not lib.drgnn_nccl_available():
            raise _lib.DrgnnError('no NCCL library could be bound (DRGNN_NCCL_LIB, libnccl.so.2)')
        dist = torch.distributed
        if id_bytes is None:
            if world == 1:
                rank = 0
                id_bytes = self.unique_id()
            else:
                if not (dist.is_available() and dist.is_initialized()):
                    raise _lib.DrgnnError('NcclComm needs the unique id of rank 0 (id_bytes) or a process group')
                world, rank = dist.get_world_size(group), dist.get_rank(group)
                box = [self.unique_id() if rank == 0 else None]
                dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0,
                                           group=group)
                id_bytes = box[0]
        if len(id_bytes) != _lib.NCCL_ID_BYTES:
            raise _lib.DrgnnError('an NCCL unique id has %d bytes' % _lib.NCCL_ID_BYTES)
        self.world, self.rank = int(world), int(rank)
        comm = _lib.VP()
        buf = C.create_string_buffer(bytes(id_bytes), _lib.NCCL_ID_BYTES)
        _lib.check(lib.drgnn_nccl_init(C.byref(comm), self.world, self.rank, C.cast(buf, _lib.VP)), 'drgnn_nccl_init')
        self._comm = comm.value

    @staticmethod
    def unique_id():
        import ctypes as C
        from . import _lib
        buf = C.create_string_buffer(_lib.NCCL_ID_BYTES)
        _lib.check(_lib.load().drgnn_nccl_unique_id(C.cast(buf, _lib.VP)), 'drgnn_nccl_unique_id')
        return buf.raw

    def all_reduce_(self, t):
        from . import _lib
        _lib.require_cuda(t)
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise _lib.DrgnnError('drgnn_nccl_allreduce sums contiguous float32 buffers')
        if self._comm is None:
            raise _lib.DrgnnError('the communicator is closed')
        _lib.call('drgnn_nccl_allreduce', self._comm, t.data_ptr(), t.numel(), _lib.stream_ptr(t.device))
        return t

    def close(self):
        from . import _lib
        if self._comm is not None:
            _lib.load().drgnn_nccl_destroy(self._comm)
            self._comm = None
