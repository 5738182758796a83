"""Where does a training step go?  Times (CUDA events, graph replay) on one GPU:
  prep only   - structure pass of one batch, repeated
  step only   - everything after the structure pass, repeated on one prepared batch
  pipelined   - Engine.train_resident (prep of i+1 on a side stream while step i computes)
  serial      - prep + step on one stream
Usage: python tools/step_breakdown.py [cfg2] [steps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402
from deeprank_gnn_b200.data import PackedBatch  # noqa: E402
from deeprank_gnn_b200.engine import Engine  # noqa: E402


def timed(fn, n):
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(n):
        fn(i)
    e.record()
    torch.cuda.synchronize()
    return 1e3 * s.elapsed_time(e) / n


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    cfg = bench.workload_config(name, None)
    _graphs, batches = bench.make_pool(cfg, 8, seed=0)
    packed = [PackedBatch.from_batch(b, idx16=True, edge_attr=cfg['net'] == 'sGAT') for b in batches]
    for graph in (True, False):
        eng = Engine(cfg['net'], cfg['feat'], 1, 1, hidden=cfg['hidden'], device='cuda:0', lr=1e-3, graph=graph, seed=0)
        ds = [eng.upload(pb, slot=i) for i, pb in enumerate(packed)]
        for d in ds:
            eng.step(d)
        eng.train_resident(ds, steps=16)
        t_prep = timed(lambda i: eng._prepare_any(ds[i % 8]), n)
        eng._prepare_any(ds[0])
        t_step = timed(lambda i: eng.step(ds[0], prepared=True), n)
        t_serial = timed(lambda i: eng.step(ds[i % 8]), n)
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        eng.train_resident(ds, steps=n)
        e.record()
        torch.cuda.synchronize()
        t_pipe = 1e3 * s.elapsed_time(e) / n
        if graph:
            import ctypes
            from deeprank_gnn_b200 import _lib
            eng.phase_timers, eng.use_graph = 2, False      # eager launches with the phase clocks on
            for _ in range(3):
                eng.step(ds[0], prepared=True)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            eng.step(ds[0], prepared=True)
            ev1.record()
            torch.cuda.synchronize()
            print('  the clocked launch, CUDA events around it: %.1f us' % (1e3 * ev0.elapsed_time(ev1)))
            eng.phase_timers, eng.use_graph = False, True
            ph = (ctypes.c_uint64 * 32)()
            _lib.check(_lib.load().drgnn_debug_phase_cycles(ph), 'phase')
            names = ['stage', 'AX', 'Z1', 'P1', 'AP', 'Z2', 'P2', 'R', 'fc1', 'fc2', 'loss+headbwd+dR', 'dZ2stage',
                     'dW2/dAP', 'dP1', 'dZ1stage', 'dW1']
            tot = ph[16] - ph[0]
            print('graph-step kernel, CTA of graph 0: %d cycles total' % tot)
            print('  ' + ' | '.join('%s %d' % (nm, ph[i + 1] - ph[i]) for i, nm in enumerate(names)))
            if ph[25] > ph[0]:
                print('  staging split: extents %d | bulk issue %d | head vectors %d | weight transposes %d | CTA barrier %d | '
                      'wait for blob + features %d | header check + cluster arrive %d'
                      % (ph[20] - ph[0], ph[21] - ph[20], ph[22] - ph[21], ph[23] - ph[22], ph[24] - ph[23],
                         ph[25] - ph[24], ph[1] - ph[25]))
            if ph[28] > ph[26] > 0:
                print('  repeated inside the launch (warm code, same data): Z1 %d | AX %d' % (ph[27] - ph[26], ph[28] - ph[27]))
            if ph[18] > ph[16]:
                print('  in-kernel reduction: grid barrier %d | reduce + Adam %d' % (ph[17] - ph[16], ph[18] - ph[17]))
            if eng._last_path == 'step3':
                _lib.check(_lib.load().drgnn_debug_phase3_cycles(ph), 'phase3')
                n3 = ['stage', 'zin1', 'Z1', 'P1', 'zin2', 'Z2', 'P2', 'readout', 'head', 'headbwd', 'dZ2', 'dW2/dzin2',
                      'dP1', 'dZ1', 'dW1', 'reduce']
                from deeprank_gnn_b200 import ops
                nct = min(2048, cfg['batch'] * ops.net_step_last()[1] * (2 if cfg['net'] == 'GINet' else 1))
                ct = (ctypes.c_uint64 * (2 * nct))()
                _lib.check(_lib.load().drgnn_debug_cta_times(ct, nct), 'cta times')
                st0 = [ct[2 * i] for i in range(nct)]
                en0 = [ct[2 * i + 1] for i in range(nct)]
                t00 = min(st0)
                dur = sorted(e - s_ for s_, e in zip(st0, en0))
                print('  CTAs (%d): start skew max %.1f us | per-graph work min / median / max %.1f / %.1f / %.1f us | last CTA done at %.1f us'
                      % (nct, (max(st0) - t00) / 1e3, dur[0] / 1e3, dur[len(dur) // 2] / 1e3, dur[-1] / 1e3, (max(en0) - t00) / 1e3))
                print('general cluster step kernel (tiles %d), block 0: %d cycles total'
                      % (ops.net_step_last()[1], max(ph[15], ph[16]) - ph[0]))
                print('  ' + ' | '.join('%s %d' % (nm, ph[i + 1] - ph[i]) for i, nm in enumerate(n3)
                                        if ph[i + 1] >= ph[i]))
            _lib.check(_lib.load().drgnn_debug_blob_cycles(ph), 'bphase')
            bnames = ['load+minmax', 'relabel', 'scatter', 'count+scan', 'emit']
            print('blob structure kernel, CTA of graph 0: %d cycles total' % (ph[5] - ph[0]))
            print('  ' + ' | '.join('%s %d' % (nm, ph[i + 1] - ph[i]) for i, nm in enumerate(bnames)))
            if ph[9] and ph[11] and ph[5] > ph[0]:
                print('  emit split: level-0 CSR (slots + rank sweep) %d | member lists + pointers %d | pooled rows / columns %d | '
                      'rest (weights, first aggregation, header) %d' % (ph[9] - ph[4], ph[10] - ph[9], ph[11] - ph[10], ph[5] - ph[11]))
            if ph[6] and ph[7]:
                print('  emit split (sGAT weights): lists %d | pooled sums %d | CSC order + header %d'
                      % (ph[6] - ph[4], ph[7] - ph[6], ph[5] - ph[7]))
            _lib.check(_lib.load().drgnn_debug_structure_cycles(ph), 'sphase')
            snames = ['edges', 'CSR', 'CSC', 'relabel', 'members', 'coarsen', 'CSC1', 'level1']
            print('structure kernel, CTA of graph 0: %d cycles total' % (ph[8] - ph[0]))
            print('  ' + ' | '.join('%s %d' % (nm, ph[i + 1] - ph[i]) for i, nm in enumerate(snames)))
        print('%s graph=%s: prep %.1f us | step (after prep) %.1f us | serial %.1f us | pipelined %.1f us'
              % (name, graph, t_prep, t_step, t_serial, t_pipe))


if __name__ == '__main__':
    main()
