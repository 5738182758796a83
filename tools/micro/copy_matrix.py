"""Which factor changes the H2D rate of the feeder's copies?  (source buffers x engine present)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from deeprank_gnn_b200.data import PackedBatch
cfg = bench.workload_config('cfg2', None)
_g, batches = bench.make_pool(cfg, 16, seed=0)
packed = [PackedBatch.from_batch(b, idx16=True, edge_attr=False) for b in batches]
numel = packed[0].numel
plain = [torch.empty(numel, dtype=torch.float32).pin_memory() for _ in range(16)]
filled = [torch.randn(numel, dtype=torch.float32).pin_memory() for _ in range(16)]
zeros = [torch.zeros(numel, dtype=torch.float32, pin_memory=True) for _ in range(16)]
dev = [torch.empty(packed[0].capacity_numel, dtype=torch.float32, device='cuda') for _ in range(4)]
devx = [torch.empty(numel, dtype=torch.float32, device='cuda') for _ in range(4)]
cs = torch.cuda.Stream()
n = 400


def rate(src, name, sliced=True):
    for rep in range(2):
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(cs)
        with torch.cuda.stream(cs):
            for i in range(n):
                if sliced:
                    dev[i % 4][:src[i % 16].numel()].copy_(src[i % 16], non_blocking=True)
                else:
                    devx[i % 4].copy_(src[i % 16], non_blocking=True)
        e.record(cs)
        torch.cuda.synchronize()
    us = 1e3 * s.elapsed_time(e) / n
    print('%-44s %.1f us = %.1f GB/s' % (name, us, 4 * numel / us / 1e3))


print('numel', numel, 'bytes', 4 * numel, 'data_ptr % 4096 of a record:', packed[0].buf.data_ptr() % 4096)
rate(plain, 'empty().pin_memory()')
rate(plain, 'empty().pin_memory(), whole destination', sliced=False)
rate(filled, 'randn().pin_memory()')
rate(zeros, 'zeros(pin_memory=True)')
rate([pb.buf for pb in packed], 'PackedBatch.buf')
repinned = [torch.empty(pb.numel, dtype=torch.float32).copy_(pb.buf).pin_memory() for pb in packed]
rate(repinned, 'PackedBatch.buf copied to pageable, .pin_memory()')
zf = [torch.zeros(numel, dtype=torch.float32, pin_memory=True).normal_() for _ in range(16)]
rate(zf, 'zeros(pin_memory=True).normal_()')
ef = [torch.empty(numel, dtype=torch.float32, pin_memory=True).zero_() for _ in range(16)]
rate(ef, 'empty(pin_memory=True).zero_()')
ez = [torch.zeros(numel, dtype=torch.float32).pin_memory() for _ in range(16)]
rate(ez, 'zeros().pin_memory()  (all-zero content)')
from deeprank_gnn_b200.engine import Engine
eng = Engine('GINet', 32, 1, 1, hidden=cfg['hidden'], device='cuda:0', lr=1e-3, graph=True, seed=0)
rate(plain, 'engine created: empty().pin_memory()')
rate([pb.buf for pb in packed], 'engine created: PackedBatch.buf')
ds = [eng.upload(pb, slot=i) for i, pb in enumerate(packed)]
eng.train_resident(ds, steps=64)
rate(plain, 'engine used: empty().pin_memory()')
rate([pb.buf for pb in packed], 'engine used: PackedBatch.buf')
