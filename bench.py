#!/usr/bin/env python
"""Benchmark of the hot path: GINet training step (structure pass + forward + backward +
[all-reduce] + Adam) on synthetic protein-interface residue graphs.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2]

Prints ONE JSON line (rank 0).  Definitions (DESIGN.md "Measurement"):
  value     graphs/s over all ranks, inputs resident in HBM before the timed region; the steps
            rotate over a pool of distinct batches larger than L2, replayed as CUDA graphs
  e2e       graphs/s through Engine.train_batches(): packed batches in pinned HOST memory, one
            H2D copy per step and a D2H read of loss + predictions inside the timed region
  roofline  the aggregation kernel (gather -> segmented sum): algorithmic bytes / launch time,
            measured with CUDA events on a concatenated stream of graphs >> L2, vs the measured
            HBM copy bandwidth in MEASURED_PEAKS.json; "in_situ" = the same kernel at the B=64
            step size (launch-latency bound)
  cpu_baseline   the oracle (pure-torch restatement of the reference, which cannot be imported
            here) timed on the host cores of this box, same workload
  --impl reference   times that CPU path alone and prints its own line
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'protein-interface graphs/sec (GINet fwd+bwd, batch=64)'      # BASELINE.json (cfg2)
UNIT = 'graphs/s'
ALIGN_STEPS = 16      # untimed steps enqueued right before the start event (device-side rank alignment)


def metric_name(cfg):
    """BASELINE.json's metric for the headline workload; the same wording with the network / batch of the
    other workloads."""
    if cfg['net'] == 'GINet' and cfg['batch'] == 64:
        return METRIC
    return 'protein-interface graphs/sec (%s%s fwd+bwd, batch=%d)' % (cfg['net'], cfg.get('layers_tag', ''), cfg['batch'])


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=400)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--repeats', type=int, default=25,
                    help='the K-step timed region is measured this many times; the median is reported')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=['cfg2', 'cfg3', 'cfg4', 'cfg5'])
    ap.add_argument('--batch', type=int, default=None, help='graphs per GPU per step (default: the config batch)')
    ap.add_argument('--global-batch', type=int, default=None,
                    help='STRONG scaling (SURVEY 8e): graphs per step over ALL GPUs (cfg4: 256, cfg5: 512); each rank '
                         'takes global / gpus of them.  Default: weak scaling, --batch graphs per GPU')
    ap.add_argument('--layers', type=int, default=2, choices=[2, 3],
                    help='3: the three-layer sGAT / FoutNet throughput variant (BASELINE config 3 "sGAT 3-layer")')
    ap.add_argument('--pool', type=int, default=64, help='distinct batches rotated through (must exceed L2)')
    ap.add_argument('--no-graph', action='store_true', help='launch kernels eagerly instead of CUDA-graph replay')
    ap.add_argument('--no-roofline', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--stream-nodes', type=int, default=6553600, help='nodes of the roofline stream (>> L2)')
    args = ap.parse_args()
    if args.global_batch is not None:
        if args.batch is not None:
            ap.error('--batch (per GPU, weak scaling) and --global-batch (strong scaling) exclude each other')
        if args.global_batch % max(1, args.gpus) or args.global_batch < args.gpus:
            ap.error('--global-batch must be a positive multiple of --gpus')
        args.batch = args.global_batch // args.gpus
        args.scaling = 'strong'
    else:
        args.scaling = 'weak'
    return args


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


# ------------------------------------------------------------------------------ workload
def workload_config(name, batch):
    from deeprank_gnn_b200 import synthetic
    c = dict(synthetic.CONFIGS[name])
    if batch is not None:
        c['batch'] = batch
    elif name == 'cfg4':
        c['batch'] = 32            # 256 over 8 GPUs
    elif name == 'cfg5':
        c['batch'] = 64            # 512 over 8 GPUs
    return c


def make_pool(cfg, pool, seed):
    """``pool`` distinct batches, each a random draw of ``batch`` graphs from a set of unique
    synthetic graphs (every batch is a distinct block of memory)."""
    import numpy as np
    from deeprank_gnn_b200 import synthetic
    from deeprank_gnn_b200.data import Batch
    B = cfg['batch']
    n_unique = max(B, min(B * pool, 1024))
    graphs = synthetic.make_graphs(cfg, count=n_unique, seed=seed, internal=False)
    rng = np.random.default_rng(seed)
    batches = []
    for _ in range(pool):
        idx = rng.choice(n_unique, size=B, replace=False)
        batches.append(Batch.from_data_list([graphs[i] for i in idx]))
    return graphs, batches


# ------------------------------------------------------------------------------ clocks
class ClockSampler(object):
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.FIELDS,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            parts = [p.strip() for p in r.split(',')]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'),
                                 parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------ CPU arm
def cpu_steps(cfg, batches, seconds, steps=None, warmup=1):
    """Time the oracle (CPU restatement of the reference step) on `batches`.  Returns
    (graphs/s, ms/step, steps run)."""
    import copy
    from oracle import nets as onets
    from oracle import pyg_min
    from oracle import step as ostep
    torch.set_num_threads(os.cpu_count() or 1)
    net = {'GINet': onets.GINet, 'sGAT': onets.sGAT, 'FoutNet': onets.FoutNet}[cfg['net']]
    if cfg.get('layers_tag'):
        if cfg['net'] != 'sGAT':
            raise SystemExit('--layers 3 has a CPU oracle for sGAT only')
        net = onets.sGAT3
    onets.LITERAL = cfg['net'] != 'FoutNet'       # the literal per-node Fout loop takes seconds per batch
    torch.manual_seed(0)
    model = net(cfg['feat'], 1, 1, hidden=cfg['hidden']).train()
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    loss_fn = ostep.make_loss('reg')

    def to_oracle(b):
        ob = pyg_min.Batch()
        for k in ('x', 'edge_index', 'edge_attr', 'y', 'pos', 'batch', 'cluster0', 'cluster1'):
            setattr(ob, k, getattr(b, k).clone())
        return ob

    n = 0
    t_total = 0.0
    i = 0
    while True:
        ob = to_oracle(batches[i % len(batches)])       # collation excluded from the timed region
        t0 = time.perf_counter()
        ostep.train_step(model, opt, loss_fn, ob)
        dt = time.perf_counter() - t0
        i += 1
        if i <= warmup:
            continue
        n += 1
        t_total += dt
        if steps is not None:
            if n >= steps:
                break
        elif t_total >= seconds:
            break
    onets.LITERAL = True
    ms = 1e3 * t_total / n
    return cfg['batch'] / (ms / 1e3), ms, n


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    cfg = workload_config(args.workload, args.batch)
    if args.layers == 3:
        cfg['layers_tag'] = ' 3-layer'
    _graphs, batches = make_pool(cfg, min(args.pool, 8), seed=0)
    # a "step" is one batch; bound the run to a few minutes whatever K is
    gps, ms, n = cpu_steps(cfg, batches, seconds=None, steps=args.steps, warmup=max(1, min(args.warmup, 5))) \
        if args.steps * 0.2 < 240 else cpu_steps(cfg, batches, seconds=120.0, warmup=3)
    cores = torch.get_num_threads()
    line = {
        'impl': 'reference', 'metric': metric_name(cfg), 'value': gps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': n,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': describe(cfg, args),
        'cpu_baseline': {'value': gps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d training steps (fwd+bwd+Adam) of batch %d on %d host threads; CPU oracle = '
                                   'pure-torch restatement of the reference step (the reference itself needs '
                                   'torch_geometric / torch_scatter, absent here)' % (n, cfg['batch'], cores)},
        'e2e': {'value': gps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def describe(cfg, args):
    nodes = cfg['nodes']
    d = {'workload': '%s: synthetic residue graphs, %s nodes / %s directed edges, %d node features, %s hidden %s, '
                     'batch %d per GPU' % (args.workload, nodes, cfg.get('edges', '8 per node'), cfg['feat'], cfg['net'],
                                           tuple(cfg['hidden']), cfg['batch']),
         'path': cfg['net'] + ' train step', 'batch_per_gpu': cfg['batch'], 'global_batch': cfg['batch'] * args.gpus,
         'step': 'structure pass + forward + loss + backward + gradient reduction / exchange + Adam',
         'l2': 'inputs larger than L2: %d distinct batches rotated' % args.pool,
         'parallelism': 'dp%d' % args.gpus}
    return d


# ------------------------------------------------------------------------------ roofline
def aggregation_roofline(cfg, graphs, n_nodes_target, hbm_gbs, peak_src, in_situ_batch):
    """Time the aggregation kernel with CUDA events.  Stream: the unique graphs tiled to
    >= n_nodes_target nodes (one launch over a working set >> L2).  Algorithmic bytes =
    4NC (read each source row once) + 4NC (write) + 4E (col) + 4(N+1) (rowptr)  [SURVEY 8d]."""
    from deeprank_gnn_b200 import ops
    from deeprank_gnn_b200.data import Batch
    dev = torch.device('cuda', torch.cuda.current_device())
    C = cfg['feat']
    base = Batch.from_data_list(graphs[:256] if len(graphs) >= 256 else graphs)
    st = ops.structure_build(base._node_ptr.to(dev), base._edge_ptr.to(dev), base.edge_index.to(dev),
                             base.cluster0.to(dev), base._max_n, base._max_e)
    n0, e0 = base.x.size(0), base.edge_index.size(1)
    reps = max(1, (n_nodes_target + n0 - 1) // n0)
    N, E = n0 * reps, e0 * reps
    # replicate the CSR block-diagonally on the device
    off_n = (torch.arange(reps, device=dev, dtype=torch.int32) * n0).view(-1, 1)
    off_e = (torch.arange(reps, device=dev, dtype=torch.int32) * e0).view(-1, 1)
    rowptr = torch.cat([(st.rowptr0[:n0].view(1, -1) + off_e).reshape(-1),
                        torch.tensor([E], dtype=torch.int32, device=dev)])
    col = (st.col0[:e0].view(1, -1) + off_n).reshape(-1).contiguous()
    nb = base.num_graphs
    tile_ptr = torch.cat([(base._node_ptr[:nb].to(dev).view(1, -1) + off_n).reshape(-1),
                          torch.tensor([N], dtype=torch.int32, device=dev)])
    tile_eptr = torch.cat([(base._edge_ptr[:nb].to(dev).view(1, -1) + off_e).reshape(-1),
                           torch.tensor([E], dtype=torch.int32, device=dev)])
    x = torch.randn(N, C, device=dev)
    out = torch.empty(N, C, device=dev)
    alg_bytes = 4.0 * N * C * 2 + 4.0 * E + 4.0 * (N + 1)

    def timed(fn, iters=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / iters      # ms

    t_rows = timed(lambda: ops.aggregate(x, rowptr, col, out))
    t_tiled = timed(lambda: ops.aggregate(x, rowptr, col, out, tile_ptr=tile_ptr, tile_eptr=tile_eptr,
                                          max_tile_rows=base._max_n, max_tile_edges=base._max_e))
    best, t_best = ('aggregate_tiled_kernel', t_tiled) if t_tiled < t_rows else ('aggregate_rows_kernel', t_rows)
    achieved = alg_bytes / (t_best * 1e-3) / 1e9
    # in situ: the same kernel on ONE step-sized batch, averaged over back-to-back launches
    bb = in_situ_batch
    stb = ops.structure_build(bb._node_ptr.to(dev), bb._edge_ptr.to(dev), bb.edge_index.to(dev), bb.cluster0.to(dev),
                              bb._max_n, bb._max_e)
    xb = bb.x.to(dev)
    ob = torch.empty_like(xb)
    np_b, ep_b = bb._node_ptr.to(dev), bb._edge_ptr.to(dev)

    def graph_timed(fn, reps=50, iters=20):
        """Average launch time of `fn` replayed from a CUDA graph of `reps` back-to-back launches
        (removes the Python / ctypes launch overhead from a step-sized kernel)."""
        fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                fn()
        g.replay()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(iters):
            g.replay()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / (iters * reps)
    t_is_rows = graph_timed(lambda: ops.aggregate(xb, stb.rowptr0, stb.col0, ob))
    t_is_tiled = graph_timed(lambda: ops.aggregate(xb, stb.rowptr0, stb.col0, ob, tile_ptr=np_b, tile_eptr=ep_b,
                                                   max_tile_rows=bb._max_n, max_tile_edges=bb._max_e))
    t_is = min(t_is_rows, t_is_tiled)
    nb_, eb_ = xb.size(0), bb.edge_index.size(1)
    is_bytes = 4.0 * nb_ * C * 2 + 4.0 * eb_ + 4.0 * (nb_ + 1)
    del x, out
    return {
        'bound': 'hbm', 'kernel': best, 'achieved': achieved, 'peak': hbm_gbs, 'unit': 'GB/s',
        'frac': achieved / hbm_gbs,
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on this stream: NOT measured
        # in this run - read from the committed ncu --set full capture (996.4 MB + 791.0 MB)
        'traffic': _profile_csv_metric(os.path.join(ROOT, 'profiles', 'r1_agg_tiled_final_ncu_raw.csv'),
                                       'aggregate_tiled_kernel', ['dram__bytes_read.sum', 'dram__bytes_write.sum'])
        if (N == 6553600 and C == 32 and best == 'aggregate_tiled_kernel') else None,
        'traffic_source': 'profiles/r1_agg_tiled_final_ncu_raw.csv (ncu --set full of the same launch; not measured in this run)',
        'peak_source': peak_src,
        'algorithmic_bytes_per_launch': alg_bytes, 'launch_ms': t_best,
        'stream': {'nodes': N, 'directed_edges': E, 'channels': C, 'rows_kernel_ms': t_rows, 'tiled_kernel_ms': t_tiled,
                   'rows_kernel_gbs': alg_bytes / (t_rows * 1e-3) / 1e9,
                   'tiled_kernel_gbs': alg_bytes / (t_tiled * 1e-3) / 1e9},
        'in_situ': {'nodes': nb_, 'directed_edges': eb_, 'launch_us': t_is * 1e3, 'rows_kernel_us': t_is_rows * 1e3,
                    'tiled_kernel_us': t_is_tiled * 1e3,
                    'achieved': is_bytes / (t_is * 1e-3) / 1e9, 'frac': is_bytes / (t_is * 1e-3) / 1e9 / hbm_gbs,
                    'note': 'step-sized launch replayed from a CUDA graph (L2 resident, launch-latency bound)'},
    }



def _profile_csv_metric(path, kernel_substr, metrics):
    """Sum of `metrics` (ncu --page raw --csv columns) over the launches of the kernels matching
    `kernel_substr` in a committed profile, divided by the number of launches; None when absent."""
    import csv
    if not os.path.exists(path):
        return None
    with open(path, newline='') as f:
        rows = list(csv.reader(f))
    if len(rows) < 3:
        return None
    head = rows[0]
    try:
        kcol = head.index('Kernel Name')
        cols = [head.index(m) for m in metrics]
    except ValueError:
        return None
    tot, n = 0.0, 0
    for r in rows[2:]:
        if len(r) <= max(cols + [kcol]) or kernel_substr not in r[kcol]:
            continue
        try:
            vals = [float(r[c].replace(',', '')) for c in cols]
        except ValueError:
            continue
        units = [rows[1][c] for c in cols]
        scale = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
        tot += sum(v * scale.get(u, 1.0) for v, u in zip(vals, units))
        n += 1
    return tot / n if n else None


def step_kernel_traffic(kernel, workload):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel from the committed
    `ncu --set full` capture of this workload (profiles/); (None, reason) when there is none."""
    for name in ('r2_final3_step_%s_ncu_raw.csv' % workload, 'r2_step_%s_ncu_raw.csv' % workload,
                 'r1_cluster_step_kernels_ncu_raw.csv' if workload == 'cfg2' else ''):
        if not name:
            continue
        path = os.path.join(ROOT, 'profiles', name)
        v = _profile_csv_metric(path, kernel, ['dram__bytes_read.sum', 'dram__bytes_write.sum'])
        if v is not None:
            return v, 'profiles/' + name + ' (ncu --set full, per launch; not measured in this run)'
    return None, 'no ncu capture of this kernel committed'


def step_bytes(batch, n_nodes, n_edges, feat, B, n_par, row_floats_not_moved=0):
    """(algorithmic, implementation) bytes of one whole-step launch.

    ALGORITHMIC = SURVEY 8d's per-graph figure of the ideally fused step, from the batch's own sizes: forward 4nF (x)
    + 16e (int64 edge_index as the reference delivers it) + 4e.ne (edge_attr) + 8n (cluster0) + 8 K0 (cluster1) + 8n
    (batch) + 4.out, forward + backward = 2x  (49.2 KB per cfg2 graph -> 98.4 KB per graph and step).
    IMPLEMENTATION = what the step kernel moves through L2 per launch: feature tiles + structure blobs read once + the
    per-graph gradient rows it writes and re-reads in its reduction (head v2: without the fc1.weight part, which the
    reduction forms from the fc1.bias gradient and read-out rows) + parameters / Adam state."""
    ne = int(batch.edge_attr.size(1)) if batch.edge_attr.dim() > 1 else 1
    n_b, e_b, k0_b = int(batch.x.size(0)), int(batch.edge_index.size(1)), int(batch.cluster1.numel())
    alg = 2 * (4 * n_b * feat + 16 * e_b + 4 * e_b * ne + 8 * n_b + 8 * k0_b + 8 * n_b + 4 * B)
    blob_bytes = 4 * (32 * B + 9 * n_nodes + 5 * B + 3 * n_edges)
    row = n_par + 4 - int(row_floats_not_moved)
    impl = 4 * n_nodes * feat + blob_bytes + 2 * 4 * B * row + 3 * 4 * n_par
    return alg, impl


def dense_roofline():
    """The dense per-node transform on the tensor pipe (cfg4 shape), from the committed ncu capture."""
    path = os.path.join(ROOT, 'profiles', 'r2_dense_summary.json')
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return json.load(f)

# ------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device: the hot path has no CPU fallback')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    from deeprank_gnn_b200.parallel import bind_to_gpu_numa_node, prefer_gpu_numa_memory
    # the pinned host batches must live in the memory next to the GPU's PCIe root (H2D copies out of the other
    # socket's memory are 15-40 % slower): several ranks bind their CPUs, a single process only sets its memory
    # policy (it keeps every core for the CPU baseline and resets the policy before that leg)
    if os.environ.get('DRGNN_BENCH_NUMA', '1') == '0':
        numa = 'not bound (DRGNN_BENCH_NUMA=0)'
    else:
        numa = bind_to_gpu_numa_node(local) if world > 1 else prefer_gpu_numa_memory(local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # keep stdout to the one JSON line: NCCL prints its version banner (and anything else) to stdout
        # unless told otherwise
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=dev)
    import __graft_entry__
    native = None
    if rank == 0:
        __graft_entry__.build()
        from deeprank_gnn_b200 import build as _build
        native = _build.build_record()
    if world > 1:
        dist.barrier()
    from deeprank_gnn_b200 import _lib
    from deeprank_gnn_b200.data import PackedBatch
    from deeprank_gnn_b200.engine import Engine

    cfg = workload_config(args.workload, args.batch)
    if args.layers == 3:
        cfg['layers_tag'] = ' 3-layer'
    B = cfg['batch']
    graphs, batches = make_pool(cfg, args.pool, seed=1000 * rank)
    # compact feeder records (what NeuralNet builds): uint16 graph-local edge ids, edge attributes only for sGAT
    packed = [PackedBatch.from_batch(b, idx16=True, edge_attr=cfg['net'] == 'sGAT') for b in batches]
    eng = Engine(cfg['net'], cfg['feat'], 1, 1, hidden=cfg['hidden'], device=dev, lr=0.001, graph=not args.no_graph,
                 seed=0, layers=args.layers)
    B_global = B * world
    pool_bytes = sum(p.nbytes for p in packed)

    # ---- inputs resident in HBM: one staging slot (and one captured graph) per pool batch
    resident = [eng.upload(pb, slot=i) for i, pb in enumerate(packed)]
    k0 = _lib.kernel_count
    eng.use_graph = False
    eng.step(resident[0], B_global=B_global)           # eager: count the kernels of one step
    kernels_per_step = _lib.kernel_count - k0
    eng.use_graph = not args.no_graph
    for d in resident:                                   # capture / first-touch everything (untimed)
        eng.step(d, B_global=B_global)
    K, W, REP = args.steps, max(args.warmup, 3), max(1, args.repeats)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()                                  # nvidia-smi forks here, far away from any timed region

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def resident_region():
        """barrier + synchronize | ALIGN_STEPS untimed steps | start event | EXACTLY K steps | stop event |
        synchronize + barrier.  The untimed steps are enqueued without a host sync in front of the start event:
        with several ranks every step ends in the peer exchange, so all ranks' start events sit within one
        NVLink hop of each other whatever the skew of the hosts leaving the barrier (the host is far ahead of
        the device: one graph launch per 16 steps).  No step of the region is issued eagerly (chunk graphs of
        16 steps + one of K % 16)."""
        sync_all()
        eng.train_resident(resident, steps=ALIGN_STEPS, B_global=B_global, start=0)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        eng.train_resident(resident, steps=K, B_global=B_global, start=ALIGN_STEPS)
        ev1.record()
        sync_all()
        return ev0.elapsed_time(ev1)          # ms

    resident_region()                                    # captures the chunk graphs (untimed)
    eng.train_resident(resident, steps=W, B_global=B_global)     # the W warm-up steps
    t_res = [resident_region() for _ in range(REP)]
    final_loss = float(eng.ws.loss.item())
    eng.validate()

    # ---- end to end: pinned host batches -> train_batches (H2D + step + D2H per step)
    seq = [packed[i % len(packed)] for i in range(K)]
    eng.train_batches(seq[:max(8, W)], B_global=B_global)

    def e2e_region():
        sync_all()
        eng.train_resident(resident, steps=ALIGN_STEPS, B_global=B_global, start=0)   # untimed: aligns the ranks
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        losses, _preds = eng.train_batches(seq, B_global=B_global)     # returns after its own host sync
        e1.record()
        sync_all()
        assert bool(torch.isfinite(losses).all()), 'non-finite loss in the end-to-end run'
        return e0.elapsed_time(e1)

    e2e_region()
    t_e2e = [e2e_region() for _ in range(max(1, min(REP, 15)))]
    clocks = sampler.stop() if rank == 0 else None
    eng.validate()

    def over_ranks(ts):
        """per-repeat MAX over ranks, then the median over the repeats"""
        t = torch.tensor(ts, device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        v = sorted(t.tolist())
        return v[len(v) // 2], v[0], v[-1]
    t_dev, t_dev_min, t_dev_max = over_ranks(t_res)
    t_e2e_med, t_e2e_min, t_e2e_max = over_ranks(t_e2e)
    # data-parallel invariant: every rank holds bit-identical weights after the same steps
    weights_equal = True
    if world > 1:
        chk = eng.params.data.view(torch.int32).to(torch.int64)
        mine = torch.stack([chk.sum(), (chk * torch.arange(1, chk.numel() + 1, device=dev)).sum()])
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        weights_equal = all(bool(torch.equal(c, allc[0])) for c in allc)
        assert weights_equal, 'the weights differ between ranks after the same training steps'
    t_e2e = t_e2e_med

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    hbm, peak_src = peaks()
    line = {
        'metric': metric_name(cfg), 'value': B_global * args.steps / (t_dev * 1e-3), 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t_dev / args.steps, 'higher_is_better': True,
        'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': describe(cfg, args),
        'e2e': {'value': B_global * args.steps / (t_e2e * 1e-3), 'unit': UNIT,
                'h2d_bytes_per_step': int(pool_bytes / len(packed)), 'd2h_bytes_per_step': 4 * (4 + B),
                'ms_per_step': t_e2e / args.steps, 'repeats': len(range(max(1, min(REP, 15)))),
                'ms_per_step_min_max': [t_e2e_min / K, t_e2e_max / K],
                'api': 'Engine.train_batches(PackedBatch[...]) - pinned host batch, one H2D copy, fused step, '
                       'D2H of loss + predictions'},
        'gpu_launches': kernels_per_step * args.steps,
        'kernels_per_step': kernels_per_step,
        'cuda_graph': not args.no_graph,
        'collective': eng.collective() + ('' if eng.comm_error is None else ' (peer memory unavailable: %s)' % eng.comm_error),
        'final_loss': final_loss,
        'clocks': clocks,
        'timing': {'repeats': REP, 'statistic': 'median over repeats of (max over ranks) of the K-step region',
                   'ms_per_step_min_max': [t_dev_min / K, t_dev_max / K],
                   'region': 'barrier + synchronize | %d untimed steps (device-side rank alignment, no host sync) | '
                             'CUDA event | exactly K steps (chunk CUDA graphs, none eager) | CUDA event | synchronize + '
                             'barrier' % ALIGN_STEPS},
        'weights_equal_across_ranks': weights_equal if world > 1 else None,
        'host_numa': numa,
        'native': native,
    }
    if world == 1 and not args.no_roofline:
        agg = aggregation_roofline(cfg, graphs, args.stream_nodes, hbm, peak_src, batches[0])
        # PRIMARY entry = the kernel that dominates the timed step: the whole-step kernel, one launch per step,
        # back to back on the main stream inside the chunk graphs (the structure pass of later batches runs
        # beside it on side streams), so its average launch duration over the timed region is ms_per_step.
        # Bytes per launch: step_bytes().
        step_kernel = eng.step_kernel_name()
        compact = step_kernel == 'ginet_graph_step2_kernel' and getattr(eng, 'head_v2', False)
        alg, impl_bytes = step_bytes(batches[0], packed[0].N, packed[0].E, cfg['feat'], B, int(eng.params.numel),
                                     eng.spec.Hd * eng.spec.C2 - eng.spec.C2 if compact else 0)
        us = 1e3 * t_dev / K
        traffic, traffic_src = step_kernel_traffic(step_kernel, args.workload)
        line['roofline'] = {
            'bound': 'hbm', 'kernel': step_kernel, 'achieved': alg / (us * 1e-6) / 1e9, 'peak': hbm, 'unit': 'GB/s',
            'frac': alg / (us * 1e-6) / 1e9 / hbm, 'traffic': traffic, 'traffic_source': traffic_src,
            'peak_source': peak_src, 'algorithmic_bytes_per_launch': alg,
            'algorithmic_bytes': 'SURVEY 8d, ideally fused step: 2 x (4nF + 16e + 4e.ne + 8n + 8K0 + 8n + 4.out) per graph '
                                 '= %.1f KB per graph' % (alg / B / 1e3),
            # what THIS kernel moves through L2 per launch: feature tiles + structure blobs read once + the per-graph
            # gradient rows it writes and re-reads in its reduction + parameters / Adam state
            'implementation_bytes_per_launch': impl_bytes,
            'launch_us': us, 'launches_per_step': kernels_per_step,
            'note': 'whole-step kernel at batch %d: latency / issue bound by construction (SURVEY fact 10); the HBM-bound '
                    'stream-scale kernel of the path is reported under aggregation_stream' % B,
            'aggregation_stream': agg,
        }
        dense = dense_roofline()
        if dense is not None:
            line['roofline']['dense'] = dense
    if world == 1 and not args.no_cpu:
        prefer_gpu_numa_memory(local, enable=False)        # the CPU leg allocates under the default policy
        gps, ms, n = cpu_steps(cfg, batches[:8], seconds=args.cpu_seconds, warmup=2)
        cores = torch.get_num_threads()
        line['cpu_baseline'] = {'value': gps, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'ms_per_step': ms,
                                'sample': '%d training steps (fwd+bwd+Adam) of the same batch shape, %d host threads '
                                          '(os.cpu_count=%s); oracle = pure-torch restatement of the reference'
                                          % (n, cores, os.cpu_count())}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
