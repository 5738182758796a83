"""Data-parallel sharding of a mini-batch of independent graphs.

Graphs of a mini-batch are independent (block-diagonal ``edge_index``, per-graph clusters,
per-graph read-out), so the path shards with NO data-path collective: every rank runs the
fused step on its own graphs and the only exchange is ONE all-reduce (sum) of the flat
gradient buffer before Adam.  The loss of a rank is ``sum_local(...) / B_global`` so the
summed gradients equal the gradient of the global mean whatever the shard sizes.
"""
import os

import torch


def graph_cost(d):
    """Work estimate of one graph: nodes + directed edges."""
    return int(d.num_nodes) + int(d.num_edges)


def shard_indices(costs, world, balance=True):
    """Partition graph indices ``0..len(costs)-1`` into ``world`` shards.

    ``balance=False``: contiguous equal-count chunks (fixed-size configurations).
    ``balance=True``: longest-processing-time bin packing on ``costs`` with shard sizes kept
    within one graph of each other (mixed-size batches, BASELINE config 5).  Deterministic."""
    n = len(costs)
    if world <= 0:
        raise ValueError('world must be positive')
    if not balance:
        base, rem = divmod(n, world)
        out, s = [], 0
        for r in range(world):
            k = base + (1 if r < rem else 0)
            out.append(list(range(s, s + k)))
            s += k
        return out
    cap = (n + world - 1) // world
    order = sorted(range(n), key=lambda i: (-costs[i], i))
    loads = [0] * world
    out = [[] for _ in range(world)]
    for i in order:
        r = min((r for r in range(world) if len(out[r]) < cap), key=lambda r: (loads[r], r))
        out[r].append(i)
        loads[r] += costs[i]
    return [sorted(s) for s in out]


def shard_graphs(graphs, world, rank, balance=True):
    parts = shard_indices([graph_cost(g) for g in graphs], world, balance)
    return [graphs[i] for i in parts[rank]]


def init_distributed(backend=None):
    """One process per GPU (torchrun env: RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).
    Returns (rank, world, local_rank).  NCCL on CUDA, gloo otherwise."""
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if world > 1 and not torch.distributed.is_initialized():
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        os.environ.setdefault('MASTER_PORT', '29500')
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
            torch.distributed.init_process_group(backend, device_id=torch.device('cuda', local))
        else:
            torch.distributed.init_process_group(backend)
    return rank, world, local


def all_reduce_flat_(flat, group=None):
    """The path's only collective: sum the flat gradient buffer over ranks, in place."""
    if torch.distributed.is_available() and torch.distributed.is_initialized() and \
            torch.distributed.get_world_size(group) > 1:
        torch.distributed.all_reduce(flat, group=group)
    return flat
