"""Launch the aggregation kernels a few times on a concatenated stream of synthetic graphs
(working set >> L2) - the target of `ncu --set full` captures (profiles/).  Usage:
    python tools/agg_stream.py [nodes] [channels] [launches]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from deeprank_gnn_b200 import ops, synthetic  # noqa: E402
from deeprank_gnn_b200.data import Batch  # noqa: E402


def build_stream(n_nodes, C, dev, seed=0):
    graphs = synthetic.make_graphs(dict(nodes=200, edges=1000, feat=C), count=256, seed=seed, internal=False)
    base = Batch.from_data_list(graphs)
    st = ops.structure_build(base._node_ptr.to(dev), base._edge_ptr.to(dev), base.edge_index.to(dev),
                             base.cluster0.to(dev), base._max_n, base._max_e)
    n0, e0 = base.x.size(0), base.edge_index.size(1)
    reps = max(1, (n_nodes + n0 - 1) // n0)
    N, E = n0 * reps, e0 * reps
    off_n = (torch.arange(reps, device=dev, dtype=torch.int32) * n0).view(-1, 1)
    off_e = (torch.arange(reps, device=dev, dtype=torch.int32) * e0).view(-1, 1)
    rowptr = torch.cat([(st.rowptr0[:n0].view(1, -1) + off_e).reshape(-1), torch.tensor([E], dtype=torch.int32, device=dev)])
    col = (st.col0[:e0].view(1, -1) + off_n).reshape(-1).contiguous()
    tile_ptr = torch.cat([(base._node_ptr[:-1].to(dev).view(1, -1) + off_n).reshape(-1),
                          torch.tensor([N], dtype=torch.int32, device=dev)])
    tile_eptr = torch.cat([(base._edge_ptr[:-1].to(dev).view(1, -1) + off_e).reshape(-1),
                           torch.tensor([E], dtype=torch.int32, device=dev)])
    return N, E, rowptr, col, (tile_ptr, tile_eptr, base._max_n, base._max_e)


def main():
    n_nodes = int(sys.argv[1]) if len(sys.argv) > 1 else 6553600
    C = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    launches = int(sys.argv[3]) if len(sys.argv) > 3 else 3
    dev = torch.device('cuda:0')
    N, E, rowptr, col, tiles = build_stream(n_nodes, C, dev)
    x = torch.randn(N, C, device=dev)
    out = torch.empty(N, C, device=dev)
    alg = 8.0 * N * C + 4.0 * E + 4.0 * (N + 1)
    for name, kw in (('rows', {}), ('tiled', dict(tile_ptr=tiles[0], tile_eptr=tiles[1], max_tile_rows=tiles[2],
                                                   max_tile_edges=tiles[3]))):
        ops.aggregate(x, rowptr, col, out, **kw)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(launches):
            ops.aggregate(x, rowptr, col, out, **kw)
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / launches
        print('%s: N=%d E=%d C=%d  %.3f ms  %.1f GB/s algorithmic (%.3f GB)' % (name, N, E, C, ms, alg / ms / 1e6, alg / 1e9))


if __name__ == '__main__':
    main()
