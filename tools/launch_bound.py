"""Is the resident training loop bound by the GPU or by host-side issue?  For pool sizes P:
host wall time to ISSUE n steps of Engine.train_resident (no sync) vs CUDA-event time of the same.
Usage: python tools/launch_bound.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from deeprank_gnn_b200.data import PackedBatch  # noqa: E402
from deeprank_gnn_b200.engine import Engine  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    cfg = bench.workload_config('cfg2', None)
    for pool in (2, 8, 64):
        _g, batches = bench.make_pool(cfg, pool, seed=0)
        packed = [PackedBatch.from_batch(b) for b in batches]
        eng = Engine(cfg['net'], cfg['feat'], 1, 1, hidden=cfg['hidden'], device='cuda:0', lr=1e-3, graph=True, seed=0)
        ds = [eng.upload(pb, slot=i) for i, pb in enumerate(packed)]
        for d in ds:
            eng.step(d)
        eng.train_resident(ds, steps=max(32, pool))      # captures every chunk graph of the rotation
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        s.record()
        eng.train_resident(ds, steps=n)
        e.record()
        t_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        t_wall = time.perf_counter() - t0
        print('pool %2d: host issue %.1f us/step | wall %.1f us/step | device %.1f us/step'
              % (pool, 1e6 * t_issue / n, 1e6 * t_wall / n, 1e3 * s.elapsed_time(e) / n))


if __name__ == '__main__':
    main()
