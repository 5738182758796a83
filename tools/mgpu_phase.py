"""Phase cycles of the cluster step kernel on several GPUs (torchrun): where does the multi-GPU step go?
  python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/mgpu_phase.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from deeprank_gnn_b200 import _lib  # noqa: E402
from deeprank_gnn_b200.data import PackedBatch  # noqa: E402
from deeprank_gnn_b200.engine import Engine  # noqa: E402


def main():
    rank, local, world = int(os.environ['RANK']), int(os.environ['LOCAL_RANK']), int(os.environ['WORLD_SIZE'])
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    os.environ.setdefault('NCCL_DEBUG', 'WARN')
    dist.init_process_group('nccl', device_id=dev)
    cfg = bench.workload_config('cfg2', None)
    _g, batches = bench.make_pool(cfg, 16, seed=1000 * rank)
    packed = [PackedBatch.from_batch(b) for b in batches]
    eng = Engine(cfg['net'], cfg['feat'], 1, 1, hidden=cfg['hidden'], device=dev, lr=1e-3, graph=True, seed=0)
    ds = [eng.upload(pb, slot=i) for i, pb in enumerate(packed)]
    Bg = cfg['batch'] * world
    for d in ds:
        eng.step(d, B_global=Bg)
    eng.train_resident(ds, steps=32, B_global=Bg)
    torch.cuda.synchronize()
    dist.barrier()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    eng.train_resident(ds, steps=320, B_global=Bg)
    e.record()
    torch.cuda.synchronize()
    ph = (ctypes.c_uint64 * 32)()
    _lib.check(_lib.load().drgnn_debug_phase_cycles(ph), 'phase')
    eng.validate()
    print('rank %d: %.1f us/step | kernel body %d cycles | grid barrier %d | reduce+exchange+Adam %d | %s'
          % (rank, 1e3 * s.elapsed_time(e) / 320, ph[16] - ph[0], ph[17] - ph[16], ph[18] - ph[17], eng.collective()), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
