import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from deeprank_gnn_b200.data import PackedBatch
from deeprank_gnn_b200.engine import Engine
cfg = bench.workload_config('cfg2', None)
_g, batches = bench.make_pool(cfg, 8, seed=0)
packed = [PackedBatch.from_batch(b, idx16=True, edge_attr=False) for b in batches]
eng = Engine('GINet', 32, 1, 1, hidden=cfg['hidden'], device='cuda:0', lr=1e-3, graph=True, seed=0)
ds = [eng.upload(pb, slot=i) for i, pb in enumerate(packed)]
# capture a small graph by hand with debug mode
eng.train_resident(ds, steps=8)
torch.cuda.synchronize()
print('pdl', eng._last_chunk_pdl)
use_graph, eng.use_graph = eng.use_graph, False
g = torch.cuda.CUDAGraph()
g.enable_debug_mode()
with torch.cuda.graph(g):
    for i in range(3):
        eng.step(ds[i], prepared=True)
        eng.prepare(ds[(i + 2) % 8], dependent=True)
g.debug_dump('gpurun_out/e7_chunk.dot')
eng.use_graph = use_graph
for it in range(3):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for k in range(50):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    print('3 steps + 3 dependent preps per replay: %.2f us per step' % (1e3 * s.elapsed_time(e) / 150))
