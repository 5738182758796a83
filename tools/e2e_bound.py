"""Is Engine.train_batches (pinned host batches -> H2D -> structure pass -> step -> D2H) bound by the host issue
rate, by PCIe or by the GPU?  Prints host issue time, device time and the H2D rate.
Usage: python tools/e2e_bound.py [steps]"""
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
    _g, batches = bench.make_pool(cfg, 64, seed=0)
    packed = [PackedBatch.from_batch(b, idx16=True, edge_attr=False) for b in batches]
    eng = Engine(cfg['net'], cfg['feat'], 1, 1, hidden=cfg['hidden'], device='cuda:0', lr=1e-3, graph=True, seed=0)
    seq = [packed[i % 64] for i in range(n)]
    eng.train_batches(seq[:8])
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    s.record()
    eng.train_batches(seq)
    e.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    dev = 1e3 * s.elapsed_time(e) / n
    nbytes = packed[0].nbytes
    print('host issue per step (C loop): %s us' % ('%.1f' % eng.feed_issue_us if eng.feed_issue_us else 'n/a (Python loop)'))
    print('train_batches: wall (incl. final sync) %.1f us/step | device %.1f us/step | H2D %.2f MB/step = %.1f GB/s at that rate'
          % (1e6 * (t1 - t0) / n, dev, nbytes / 1e6, nbytes / dev / 1e3))
    eng.train_batches(seq[:8], train=False)
    torch.cuda.synchronize()
    s.record()
    eng.train_batches(seq, train=False)
    e.record()
    torch.cuda.synchronize()
    sc = 1e3 * s.elapsed_time(e) / n
    print('scoring (forward + loss, no gradient): %.1f us/step = %.2f M graphs/s end to end'
          % (sc, cfg['batch'] / sc))
    # raw H2D rate of the same buffers, nothing else
    stage = torch.empty(packed[0].capacity_numel, dtype=torch.float32, device='cuda:0')
    torch.cuda.synchronize()
    s.record()
    for pb in seq:
        stage[:pb.numel].copy_(pb.buf, non_blocking=True)
    e.record()
    torch.cuda.synchronize()
    cp = 1e3 * s.elapsed_time(e) / n
    print('H2D copies alone: %.1f us/step = %.1f GB/s' % (cp, nbytes / cp / 1e3))


if __name__ == '__main__':
    main()
