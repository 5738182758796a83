"""Do H2D copies slow down while the training kernels run?  400 copies of 1.8 MB on a side stream, alone and
concurrently with Engine.train_resident on the main stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from deeprank_gnn_b200.data import PackedBatch
from deeprank_gnn_b200.engine import Engine
cfg = bench.workload_config('cfg2', None)
_g, batches = bench.make_pool(cfg, 16, seed=0)
packed = [PackedBatch.from_batch(b, idx16=True, edge_attr=False) for b in batches]
eng = Engine('GINet', 32, 1, 1, hidden=cfg['hidden'], device='cuda:0', lr=1e-3, graph=True, seed=0)
ds = [eng.upload(pb, slot=i) for i, pb in enumerate(packed)]
eng.train_resident(ds, steps=64)
torch.cuda.synchronize()
n = 400
dev = [torch.empty(packed[0].capacity_numel, dtype=torch.float32, device='cuda') for _ in range(4)]
cs = torch.cuda.Stream()


def copies():
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(cs)
    with torch.cuda.stream(cs):
        for i in range(n):
            pb = packed[i % 16]
            dev[i % 4][:pb.numel].copy_(pb.buf, non_blocking=True)
    e.record(cs)
    return s, e


for mode in ('alone', 'with kernels', 'alone', 'with kernels'):
    torch.cuda.synchronize()
    if mode == 'with kernels':
        eng.train_resident(ds, steps=1600)       # ~45 ms of kernels on the main stream
    s, e = copies()
    torch.cuda.synchronize()
    us = 1e3 * s.elapsed_time(e) / n
    print('%-13s %.1f us per %.2f MB batch = %.1f GB/s' % (mode, us, packed[0].nbytes / 1e6, packed[0].nbytes / us / 1e3))
