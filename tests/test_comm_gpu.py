"""The multi-GPU exchange kernel (``csrc/comm.cu``: per-graph rows -> local sum -> peers' memory ->
rank-ordered sum -> Adam, ONE launch) checked on a single device:

* protocol test: two "ranks" = two streams of one GPU with regions mapped in the same process;
* process test: two processes on the same GPU with real CUDA IPC handles and a gloo group, the
  whole ``Engine.step`` against the single-process full-batch step and the CPU oracle.
"""
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _regions(world, n_sum):
    from deeprank_gnn_b200.parallel import PeerComm
    nbytes = PeerComm.layout(world, n_sum)[-1]
    bufs = [torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device='cuda') for _ in range(world)]
    comms = [PeerComm(n_sum, rank=r, world=world, regions=[b.data_ptr() for b in bufs], timeout_s=5.0)
             for r in range(world)]
    return bufs, comms


@pytest.mark.parametrize('world', [1, 2, 4])
@pytest.mark.parametrize('use_partial', [False, True])
def test_peer_reduce_adam_protocol_on_one_device(lib, world, use_partial):
    from deeprank_gnn_b200 import ops
    n, B = 10697, 7
    n_sum = n + 4
    bufs, comms = _regions(world, n_sum)
    gen = torch.Generator().manual_seed(3)
    p0 = torch.randn(n, generator=gen)
    ref_p = torch.nn.Parameter(p0.clone())
    ref_opt = torch.optim.Adam([ref_p], lr=0.01)
    state = []
    for r in range(world):
        state.append(dict(p=p0.clone().cuda(), m=torch.zeros(n, device='cuda'), v=torch.zeros(n, device='cuda'),
                          step=torch.zeros(4, device='cuda'), grads=torch.zeros(n_sum, device='cuda'),
                          partial=torch.zeros(B, n_sum, device='cuda'), stream=torch.cuda.Stream()))
    torch.cuda.synchronize()
    for it in range(4):                                       # > 2 steps: both parities are re-used
        local = [torch.randn(B, n + 1, generator=gen) for _ in range(world)]
        sums = []
        for r in range(world):
            acc = torch.zeros(n + 1)
            for g in range(B):                                # graph order, like the kernel
                acc = acc + local[r][g]
            sums.append(acc)
        total = torch.zeros(n + 1)
        for r in range(world):                                # rank order, like the kernel
            total = total + sums[r]
        for r in range(world):
            st = state[r]
            if use_partial:
                st['partial'].zero_()
                st['partial'][:, :n + 1].copy_(local[r])
            else:
                st['grads'].zero_()
                st['grads'][:n + 1].copy_(sums[r])
        torch.cuda.synchronize()
        for r in range(world):
            st = state[r]
            with torch.cuda.stream(st['stream']):
                ops.peer_reduce_adam(comms[r], st['grads'], n, n_sum, partial=st['partial'] if use_partial else None, B=B,
                                     step_dev=st['step'],
                                     adam=dict(p=st['p'], m=st['m'], v=st['v'], lr=0.01, beta1=0.9, beta2=0.999, eps=1e-8))
        torch.cuda.synchronize()
        ref_opt.zero_grad()
        ref_p.grad = total[:n].clone()
        ref_opt.step()
        for r in range(world):
            st = state[r]
            assert comms[r].status() == 0
            assert torch.equal(st['grads'][:n + 1].cpu(), total), 'rank %d step %d: summed gradients' % (r, it)
            assert float(st['step'][0]) == it + 1
            assert torch.equal(st['p'], state[0]['p']), 'weights differ between ranks'
            err = float((st['p'].cpu() - ref_p.detach()).abs().max())
            assert err < 2e-6, 'Adam update differs from torch.optim.Adam: %.3e' % err


def test_peer_exchange_watchdog_reports_a_missing_peer(lib):
    from deeprank_gnn_b200 import ops
    from deeprank_gnn_b200.parallel import PeerComm
    n_sum = 516
    nbytes = PeerComm.layout(2, n_sum)[-1]
    bufs = [torch.zeros((nbytes + 3) // 4, dtype=torch.int32, device='cuda') for _ in range(2)]
    comm = PeerComm(n_sum, rank=0, world=2, regions=[b.data_ptr() for b in bufs], timeout_s=0.05)
    grads = torch.ones(n_sum, device='cuda')
    ops.peer_reduce_adam(comm, grads, 512, n_sum)            # rank 1 never runs
    assert comm.status() == 1


WORKER = r'''
import copy, os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import torch, torch.distributed as dist
from deeprank_gnn_b200 import synthetic, parallel
from deeprank_gnn_b200.data import Batch
from deeprank_gnn_b200.engine import DeviceBatch, Engine
os.environ['DRGNN_PEER_TIMEOUT_S'] = '30'
rank, world, _ = parallel.init_distributed('gloo')        # two processes, ONE device: gloo for the plumbing
torch.cuda.set_device(0)
graphs = synthetic.make_graphs('cfg2', count=8, seed=11)
full = Engine('GINet', 32, 1, 1, device='cuda:0', seed=5, peer_comm=False)
full.world = 1
full.eval()
dense = DeviceBatch.from_batch(Batch.from_data_list(graphs), 'cuda:0')
eng = Engine('GINet', 32, 1, 1, device='cuda:0', seed=5, graph=%(graph)s).eval()
assert eng.comm is not None, eng.comm_error
mine = parallel.shard_graphs(graphs, world, rank, balance=False)
if %(graph)s:
    from deeprank_gnn_b200.data import PackedBatch
    d = eng.upload(PackedBatch.from_batch(Batch.from_data_list(mine)), slot=0)
else:
    d = DeviceBatch.from_batch(Batch.from_data_list(mine), 'cuda:0')
for it in range(3):
    full.step(dense)
    loss, pred = eng.step(d, B_global=len(graphs))
    eng.validate()
    torch.cuda.synchronize()
    assert abs(float(loss[0]) - float(full.ws.loss[0])) < 1e-5 * max(1.0, abs(float(full.ws.loss[0]))), (float(loss[0]), float(full.ws.loss[0]))
    for (n, a), (_, b) in zip(eng.state_dict().items(), full.state_dict().items()):
        err = float((a - b).abs().max())
        assert err < 5e-5, (it, n, err)
flat = eng.params.data.cpu()
both = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(both, flat)
assert torch.equal(both[0], both[1]), 'weights differ between ranks'
sys.stdout.write('rank ' + str(rank) + ' ok ' + eng.collective() + chr(10)); sys.stdout.flush()
dist.destroy_process_group()
'''


@pytest.mark.parametrize('graph', [False, True])
def test_two_processes_share_gradients_through_ipc_peer_memory(lib, tmp_path, graph):
    """world_size 2 on ONE GPU: real CUDA IPC mapping, Engine.step with the fused exchange equals the
    single-process step on the full batch (the two kernels are time-sliced by the driver, so the
    in-kernel wait is also exercised across a context switch)."""
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT, 'graph': 'True' if graph else 'False'})
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29641' if graph else '29640', str(script)],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('peer-memory exchange') == 2, r.stdout


def test_nccl_bridge_single_rank_communicator(lib):
    """drgnn_nccl_* on a world of ONE (what a 1-GPU box can run): the library binds NCCL at run time, creates
    a communicator from its own unique id, and the in-place fp32 sum over one rank is the identity - eagerly
    and replayed from a CUDA graph (the step graphs capture the call)."""
    from deeprank_gnn_b200.parallel import NcclComm
    assert lib.drgnn_nccl_available() == 1
    comm = NcclComm(world=1)
    t = torch.arange(10701, device='cuda', dtype=torch.float32) * 0.25
    ref = t.clone()
    comm.all_reduce_(t)
    torch.cuda.synchronize()
    assert torch.equal(t, ref)
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            comm.all_reduce_(t)
            t.mul_(2.0)
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(t, ref * 2)
    with pytest.raises(Exception):
        comm.all_reduce_(t.double())
    comm.close()
    with pytest.raises(Exception):
        comm.all_reduce_(t)


NCCL_WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
os.environ['DRGNN_PEER_COMM'] = '0'          # no peer memory: the step falls back to the one NCCL all-reduce
os.environ['DRGNN_NCCL_NATIVE'] = '1'        # ... through the C-ABI (drgnn_nccl_allreduce)
import torch, torch.distributed as dist
from deeprank_gnn_b200 import synthetic, parallel
from deeprank_gnn_b200.data import Batch
from deeprank_gnn_b200.engine import DeviceBatch, Engine
rank, world, local = parallel.init_distributed('gloo')    # gloo carries the unique id; NCCL belongs to libdrgnn
torch.cuda.set_device(local)
dev = 'cuda:%%d' %% local
graphs = synthetic.make_graphs('cfg2', count=8, seed=11)
full = Engine('GINet', 32, 1, 1, device=dev, seed=5, peer_comm=False)
full.world = 1
full.eval()
dense = DeviceBatch.from_batch(Batch.from_data_list(graphs), dev)
eng = Engine('GINet', 32, 1, 1, device=dev, seed=5, graph=False).eval()
assert eng.comm is None and eng.nccl is not None
mine = parallel.shard_graphs(graphs, world, rank, balance=False)
d = DeviceBatch.from_batch(Batch.from_data_list(mine), dev)
for it in range(3):
    full.step(dense)
    loss, pred = eng.step(d, B_global=len(graphs))
    eng.validate()
    torch.cuda.synchronize()
    assert abs(float(loss[0]) - float(full.ws.loss[0])) < 1e-5 * max(1.0, abs(float(full.ws.loss[0])))
    for (n, a), (_, b) in zip(eng.state_dict().items(), full.state_dict().items()):
        err = float((a - b).abs().max())
        assert err < 5e-5, (it, n, err)
flat = eng.params.data.cpu()
both = [torch.zeros_like(flat) for _ in range(world)]
dist.all_gather(both, flat)
assert torch.equal(both[0], both[1]), 'weights differ between ranks'
sys.stdout.write('rank ' + str(rank) + ' ok ' + eng.collective() + chr(10)); sys.stdout.flush()
eng.nccl.close()
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='NCCL needs one device per rank')
def test_two_ranks_all_reduce_through_the_nccl_bridge(lib, tmp_path):
    """world_size 2 on TWO GPUs without peer memory: Engine.step with drgnn_nccl_allreduce equals the
    single-process step on the full batch; weights bit-identical across the ranks."""
    script = tmp_path / 'nccl_worker.py'
    script.write_text(NCCL_WORKER % {'root': ROOT})
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2',
                        '--master-addr', '127.0.0.1', '--master-port', '29642', str(script)],
                       capture_output=True, text=True, timeout=240)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count('drgnn_nccl_allreduce') == 2, r.stdout
