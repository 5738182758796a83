"""In-situ duration of every C-ABI call of one training step (eager launches, a CUDA event pair
around each call, L2-warm like the real step - unlike ncu, which flushes caches per kernel).
Usage: python tools/kernel_times.py [cfg2] [reps]"""
import collections
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from deeprank_gnn_b200 import _lib, ops  # noqa: E402
from deeprank_gnn_b200.data import PackedBatch  # noqa: E402
from deeprank_gnn_b200.engine import Engine  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    cfg = bench.workload_config(name, None)
    _g, batches = bench.make_pool(cfg, 4, seed=0)
    packed = [PackedBatch.from_batch(b) for b in batches]
    eng = Engine(cfg['net'], cfg['feat'], 1, 1, hidden=cfg['hidden'], device='cuda:0', lr=1e-3, graph=False, seed=0)
    ds = [eng.upload(pb, slot=i) for i, pb in enumerate(packed)]
    for d in ds:
        eng.step(d)
    torch.cuda.synchronize()
    records = []
    real_call = ops.call

    def timed_call(fname, *args):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        real_call(fname, *args)
        e.record()
        records.append((fname, s, e))
    ops.call = timed_call
    per_step = []
    for r in range(reps):
        records.clear()
        eng.step(ds[r % len(ds)])
        torch.cuda.synchronize()
        per_step.append([(n, 1e3 * s.elapsed_time(e)) for n, s, e in records])
    ops.call = real_call
    n = len(per_step[0])
    tot = 0.0
    for i in range(n):
        vals = sorted(st[i][1] for st in per_step)
        med = vals[len(vals) // 2]
        tot += med
        print('%2d %-28s %7.2f us' % (i, per_step[0][i][0], med))
    print('sum of medians %.1f us over %d calls' % (tot, n))


if __name__ == '__main__':
    main()
