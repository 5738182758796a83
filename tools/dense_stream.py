"""The dense per-node transform Y = relu(X W + b) on a stream of node rows >> L2, in the three math modes of
``drgnn_linear`` (0 = fp32 FMA, 1 = 3xTF32 mma.sync, 2 = 3xTF32 tcgen05 + TMEM) - the target of the
``ncu --set full`` captures behind bench.py's ``roofline.dense`` entry (profiles/r2_dense_*).  Usage:
    python tools/dense_stream.py [rows] [Fin] [Fout] [launches] [modes e.g. 012]
Prints one JSON line: per mode ms, algorithmic GB/s (4 rows (Fin + Fout) bytes) and logical TFLOP/s (2 rows Fin Fout)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deeprank_gnn_b200 import ops  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 8192000        # cfg4: 256 graphs x 500 nodes x 64 batches
    Fin = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    Fout = int(sys.argv[3]) if len(sys.argv) > 3 else 64
    launches = int(sys.argv[4]) if len(sys.argv) > 4 else 5
    modes = [int(c) for c in (sys.argv[5] if len(sys.argv) > 5 else '012')]
    dev = torch.device('cuda:0')
    g = torch.Generator(device='cuda').manual_seed(0)
    X = torch.randn(rows, Fin, device=dev, generator=g)
    W = (torch.randn(Fout, Fin, device=dev, generator=g) / Fin ** 0.5).contiguous()
    b = torch.randn(Fout, device=dev, generator=g)
    Y = torch.empty(rows, Fout, device=dev)
    ref = None
    out = {'rows': rows, 'Fin': Fin, 'Fout': Fout, 'algorithmic_bytes': 4.0 * rows * (Fin + Fout),
           'logical_flop': 2.0 * rows * Fin * Fout, 'modes': {}}
    names = {0: 'linear_fma_kernel', 1: 'linear_tf32x3_kernel (mma.sync)', 2: 'linear_tcgen05_kernel'}
    for m in modes:
        ops.linear(X, W, Fin, Fout, Y, bias=b, relu=True, math=m)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(launches):
            ops.linear(X, W, Fin, Fout, Y, bias=b, relu=True, math=m)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / launches
        chk = Y[:: max(1, rows // 4096)].double()
        if ref is None:
            ref = torch.relu(X[:: max(1, rows // 4096)].double() @ W.double().t() + b.double())
        out['modes'][str(m)] = {'kernel': names[m], 'ms': ms, 'gbs': out['algorithmic_bytes'] / ms / 1e6,
                                'tflops': out['logical_flop'] / ms / 1e9, 'max_abs_err': float((chk - ref).abs().max())}
    if 2 in modes:
        import ctypes
        from deeprank_gnn_b200 import _lib
        ph = (ctypes.c_uint64 * 8)()
        torch.cuda.synchronize()
        _lib.check(_lib.load().drgnn_debug_tc5_cycles(ph), 'tc5 cycles')
        tiles = max(int(ph[6]), 1)
        out['tc5_cycles_per_tile_cta0'] = {k: round(ph[i] / tiles) for i, k in enumerate(
            ['split_store', 'fence_bar_issue', 'prefetch_issue', 'mma_wait', 'global_stores', 'closing_barrier'])}
        out['tc5_cycles_per_tile_cta0']['tmem_ld_stage'] = round(ph[7] / tiles)
    print(json.dumps(out))


if __name__ == '__main__':
    main()
