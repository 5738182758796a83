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

METRIC = 'protein-interface graphs/sec (GINet fwd+bwd, batch=64)'
UNIT = 'graphs/s'


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=400)
    ap.add_argument('--warmup', type=int, default=20)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--workload', default='cfg2', choices=['cfg2', 'cfg3', 'cfg4', 'cfg5'])
    ap.add_argument('--batch', type=int, default=None, help='graphs per GPU per step (default: the config batch)')
    ap.add_argument('--pool', type=int, default=64, help='distinct batches rotated through (must exceed L2)')
    ap.add_argument('--no-graph', action='store_true', help='launch kernels eagerly instead of CUDA-graph replay')
    ap.add_argument('--no-roofline', action='store_true')
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--stream-nodes', type=int, default=6553600, help='nodes of the roofline stream (>> L2)')
    return ap.parse_args()


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
    _graphs, batches = make_pool(cfg, min(args.pool, 8), seed=0)
    # a "step" is one batch; bound the run to a few minutes whatever K is
    gps, ms, n = cpu_steps(cfg, batches, seconds=None, steps=args.steps, warmup=max(1, min(args.warmup, 5))) \
        if args.steps * 0.2 < 240 else cpu_steps(cfg, batches, seconds=120.0, warmup=3)
    cores = torch.get_num_threads()
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': gps, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': n,
        'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f32', 'data': 'synthetic',
        'config': describe(cfg, args, note='CPU oracle: pure-torch restatement of the reference step '
                                           '(the reference itself needs torch_geometric/torch_scatter, absent here)'),
        'cpu_baseline': {'value': gps, 'unit': UNIT, 'cores': cores, 'kind': 'port',
                         'sample': '%d training steps (fwd+bwd+Adam) of batch %d on %d host threads'
                                   % (n, cfg['batch'], cores)},
        'e2e': {'value': gps, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def describe(cfg, args, note=None):
    nodes = cfg['nodes']
    d = {'workload': '%s: synthetic residue graphs, %s nodes / %s directed edges, %d node features, %s hidden %s, '
                     'batch %d per GPU' % (args.workload, nodes, cfg.get('edges', '8 per node'), cfg['feat'], cfg['net'],
                                           tuple(cfg['hidden']), cfg['batch']),
         'path': cfg['net'] + ' train step', 'batch_per_gpu': cfg['batch'], 'global_batch': cfg['batch'] * args.gpus,
         'step': 'structure pass + forward + loss + backward + gradient reduction / exchange + Adam',
         'l2': 'inputs larger than L2: %d distinct batches rotated' % args.pool,
         'parallelism': 'dp%d' % args.gpus}
    if note:
        d['note'] = note
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
        # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel on this stream, from the
        # ncu --set full capture committed as profiles/r1_agg_tiled_final_ncu_raw.csv (996.4 MB + 791.0 MB)
        'traffic': 1787434496.0 if (N == 6553600 and C == 32 and best == 'aggregate_tiled_kernel') else None,
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
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        # keep stdout to the one JSON line: NCCL prints its version banner (and anything else) to stdout
        # unless told otherwise
        os.environ.setdefault('NCCL_DEBUG_FILE', '/dev/stderr')
        if os.environ.get('NCCL_DEBUG', 'VERSION').upper() == 'VERSION':
            os.environ['NCCL_DEBUG'] = 'WARN'
        dist.init_process_group('nccl', device_id=dev)
    import __graft_entry__
    if rank == 0:
        __graft_entry__.build()
    if world > 1:
        dist.barrier()
    from deeprank_gnn_b200 import _lib
    from deeprank_gnn_b200.data import PackedBatch
    from deeprank_gnn_b200.engine import Engine

    cfg = workload_config(args.workload, args.batch)
    B = cfg['batch']
    graphs, batches = make_pool(cfg, args.pool, seed=1000 * rank)
    # compact feeder records (what NeuralNet builds): uint16 graph-local edge ids, edge attributes only for sGAT
    packed = [PackedBatch.from_batch(b, idx16=True, edge_attr=cfg['net'] == 'sGAT') for b in batches]
    eng = Engine(cfg['net'], cfg['feat'], 1, 1, hidden=cfg['hidden'], device=dev, lr=0.001, graph=not args.no_graph,
                 seed=0)
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
    eng.train_resident(resident, steps=len(resident), B_global=B_global)   # captures the chunk graphs of one rotation (untimed)
    eng.train_resident(resident, steps=max(args.warmup, 3), B_global=B_global)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    # structure pass of step i+1 on a side stream while step i computes (Engine.train_resident)
    eng.train_resident(resident, steps=args.steps, B_global=B_global)
    ev1.record()
    torch.cuda.synchronize()
    t_dev = ev0.elapsed_time(ev1)          # ms
    if world > 1:
        dist.barrier()
        t = torch.tensor([t_dev], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_dev = float(t.item())
    final_loss = float(eng.ws.loss.item())
    eng.validate()

    # ---- end to end: pinned host batches -> train_batches (H2D + step + D2H per step)
    seq = [packed[i % len(packed)] for i in range(args.steps)]
    eng.train_batches(seq[:max(4, args.warmup)], B_global=B_global)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses, preds = eng.train_batches(seq, B_global=B_global)
    e1.record()
    torch.cuda.synchronize()
    t_e2e = e0.elapsed_time(e1)
    if world > 1:
        dist.barrier()
        t = torch.tensor([t_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    assert bool(torch.isfinite(losses).all()), 'non-finite loss in the end-to-end run'

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    hbm, peak_src = peaks()
    line = {
        'metric': METRIC, 'value': B_global * args.steps / (t_dev * 1e-3), 'unit': UNIT, 'n_gpus': world,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': t_dev / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': describe(cfg, args),
        'e2e': {'value': B_global * args.steps / (t_e2e * 1e-3), 'unit': UNIT,
                'h2d_bytes_per_step': int(pool_bytes / len(packed)), 'd2h_bytes_per_step': 4 * (4 + B),
                'ms_per_step': t_e2e / args.steps,
                'api': 'Engine.train_batches(PackedBatch[...]) - pinned host batch, one H2D copy, fused step, '
                       'D2H of loss + predictions'},
        'gpu_launches': kernels_per_step * args.steps,
        'kernels_per_step': kernels_per_step,
        'cuda_graph': not args.no_graph,
        'collective': eng.collective() + ('' if eng.comm_error is None else ' (peer memory unavailable: %s)' % eng.comm_error),
        'final_loss': final_loss,
        'clocks': clocks,
    }
    if world == 1 and not args.no_roofline:
        line['roofline'] = aggregation_roofline(cfg, graphs, args.stream_nodes, hbm, peak_src, batches[0])
        if cfg['net'] == 'GINet':
            # the kernel that dominates the step itself: whole-step cluster kernel, one launch per step on the main
            # stream (its duration is bounded by the step time measured above); algorithmic bytes = feature tiles +
            # structure blobs read once, per-graph gradient rows written and re-read by the in-kernel reduction
            pb0 = packed[0]
            n_par = int(eng.params.numel)
            blob_bytes = 4 * (32 * B + 9 * pb0.N + 5 * B + 3 * pb0.E)
            alg = 4 * pb0.N * cfg['feat'] + blob_bytes + 2 * 4 * B * (n_par + 4) + 3 * 4 * n_par
            us = 1e3 * t_dev / args.steps
            line['roofline']['in_step'] = {
                'kernel': 'ginet_graph_step2_kernel', 'algorithmic_bytes_per_launch': alg, 'launch_us_upper_bound': us,
                'achieved': alg / (us * 1e-6) / 1e9, 'frac': alg / (us * 1e-6) / 1e9 / hbm,
                'note': 'latency / issue bound at batch 64 (SURVEY fact 10): ncu shows 34 % issue-slot use and 3 MB of '
                        'DRAM traffic per launch (profiles/r1_cluster_step_kernels_ncu_raw.csv)'}
    if world == 1 and not args.no_cpu:
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
