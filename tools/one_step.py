"""A few eager (no CUDA graph) training steps of one configuration - the target of ncu captures:
  ncu --set full --import-source on --clock-control none -k regex:'graph_local|graph_step2' --launch-skip 4 -c 2 \
      -o gpurun_out/step python tools/one_step.py cfg2 4
Usage: python tools/one_step.py [cfg2] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from deeprank_gnn_b200.data import PackedBatch  # noqa: E402
from deeprank_gnn_b200.engine import Engine  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    cfg = bench.workload_config(name, None)
    _graphs, batches = bench.make_pool(cfg, 2, seed=0)
    packed = [PackedBatch.from_batch(b) for b in batches]
    eng = Engine(cfg['net'], cfg['feat'], 1, 1, hidden=cfg['hidden'], device='cuda:0', lr=1e-3, graph=False, seed=0)
    ds = [eng.upload(pb, slot=i) for i, pb in enumerate(packed)]
    for i in range(n):
        loss, _ = eng.step(ds[i % 2])
    torch.cuda.synchronize()
    eng.validate()
    print('loss %.6f' % float(loss))


if __name__ == '__main__':
    main()
