"""Seeded synthetic protein-interface residue graphs (SURVEY 8d / BASELINE.md 3).

Shapes follow the fixture ``tests/hdf5/1ATN_residue.hdf5``: two chains, strictly
inter-chain ("interface") edges stored in both directions the way
``HDF5DataSet.load_one_graph`` stores them (``DataSet.py:266-269``: first half i->j,
second half j->i, unsorted), every node of degree >= 1, ``edge_attr = tanh(-d/2+2)+1``
(``DataSet.py:96``), intra-chain two-level clustering with local ids per graph
(``clustering/mcl/depth_{0,1}``), ``len(cluster1) == n_unique(cluster0)``.
Graph ``g`` of a dataset uses the numpy stream ``default_rng(seed + g)``.
"""
import numpy as np
import torch

from .data import Batch, Data

# BASELINE.json configs 2-5
CONFIGS = {
    'cfg2': dict(net='GINet', batch=64, nodes=200, edges=1000, feat=32, hidden=(16, 32)),
    'cfg3': dict(net='sGAT', batch=64, nodes=200, edges=1000, feat=32, hidden=(16, 32)),
    'cfg4': dict(net='GINet', batch=256, nodes=500, edges=4000, feat=32, hidden=(32, 64)),
    'cfg5': dict(net='FoutNet', batch=512, nodes=(50, 1000), edges_per_node=8, feat=32, hidden=(16, 32)),
}


def _groups(rng, lo, hi, total, start_id, gap_prob):
    """Split ``total`` consecutive items into contiguous groups of random size in [lo, hi]."""
    ids = np.empty(total, dtype=np.int64)
    i, cid = 0, start_id
    while i < total:
        s = int(rng.integers(lo, hi + 1))
        ids[i:i + s] = cid
        i += s
        cid += 1
        if gap_prob > 0 and rng.random() < gap_prob:
            cid += 1                      # leave an unused id (MCL "last assignment wins" gaps)
    return ids, cid


def make_graph(n, e_directed, feat, seed, internal=True, gap_prob=0.02, name=None):
    rng = np.random.default_rng(seed)
    na = (n + 1) // 2
    nb = n - na
    if nb < 1:
        raise ValueError('need at least 2 nodes')
    half = e_directed // 2
    half = max(half, max(na, nb))
    half = min(half, na * nb)
    # every node gets one partner in the other chain, then fill up without replacement
    keys = np.concatenate([np.arange(na) * nb + rng.integers(0, nb, na),
                           rng.integers(0, na, nb) * nb + np.arange(nb)])
    keys = np.unique(keys)
    while keys.size < half:
        extra = rng.integers(0, na * nb, size=2 * (half - keys.size) + 8)
        keys = np.unique(np.concatenate([keys, extra]))
    if keys.size > half:
        # drop surplus edges but never one that would isolate a node
        order = rng.permutation(keys.size)
        a, b = keys // nb, keys % nb
        dega = np.bincount(a, minlength=na)
        degb = np.bincount(b, minlength=nb)
        keep = np.ones(keys.size, dtype=bool)
        surplus = keys.size - half
        for idx in order:
            if surplus == 0:
                break
            if dega[a[idx]] > 1 and degb[b[idx]] > 1:
                keep[idx] = False
                dega[a[idx]] -= 1
                degb[b[idx]] -= 1
                surplus -= 1
        keys = keys[keep]
    keys = keys[rng.permutation(keys.size)]
    src = keys // nb
    dst = na + keys % nb
    ind = np.stack([src, dst], axis=1)                          # (e/2, 2) like grp['edge_index']
    edge_index = np.vstack((ind, np.flip(ind, 1))).T            # DataSet.py:267
    dist = rng.uniform(1.5, 8.5, size=(ind.shape[0], 1))
    dist = np.vstack((dist, dist))
    edge_attr = np.tanh(-dist / 2 + 2) + 1                      # DataSet.py:96

    x = rng.standard_normal((n, feat)).astype(np.float32)
    pos = (rng.standard_normal((n, 3)) * 10).astype(np.float32)
    y = rng.uniform(0, 1, size=(1,)).astype(np.float32)

    ca, nxt = _groups(rng, 1, 7, na, 0, gap_prob)
    cb, _ = _groups(rng, 1, 7, nb, nxt, gap_prob)
    cluster0 = np.concatenate([ca, cb])
    uniq0 = np.unique(cluster0)
    k0a = np.unique(ca).size
    k0 = uniq0.size
    c1a, nxt1 = _groups(rng, 1, 5, k0a, 0, 0.0)
    c1b, _ = _groups(rng, 1, 5, k0 - k0a, nxt1, 0.0)
    cluster1 = np.concatenate([c1a, c1b])

    d = Data(x=torch.from_numpy(x),
             edge_index=torch.from_numpy(np.ascontiguousarray(edge_index)).long(),
             edge_attr=torch.from_numpy(edge_attr.astype(np.float32)),
             y=torch.from_numpy(y),
             pos=torch.from_numpy(pos))
    if internal:
        ia = np.stack([np.arange(na - 1), np.arange(1, na)], axis=1)
        ib = na + np.stack([np.arange(nb - 1), np.arange(1, nb)], axis=1) if nb > 1 else np.zeros((0, 2), np.int64)
        iind = np.concatenate([ia, ib]).astype(np.int64)
        d.internal_edge_index = torch.from_numpy(np.ascontiguousarray(np.vstack((iind, np.flip(iind, 1))).T)).long()
        idist = rng.uniform(1.0, 3.0, size=(iind.shape[0], 1))
        idist = np.vstack((idist, idist))
        d.internal_edge_attr = torch.from_numpy((np.tanh(-idist / 2 + 2) + 1).astype(np.float32))
    d.mol = name if name is not None else 'synth_%d' % seed
    d.cluster0 = torch.from_numpy(cluster0)
    d.cluster1 = torch.from_numpy(cluster1)
    return d


def make_graphs(cfg, count=None, seed=0, internal=True):
    """List of ``Data`` for a named BASELINE config (or a dict with the same keys)."""
    c = CONFIGS[cfg] if isinstance(cfg, str) else cfg
    count = c['batch'] if count is None else count
    out = []
    size_rng = np.random.default_rng(seed + 10 ** 6)
    for g in range(count):
        nodes = c['nodes']
        if isinstance(nodes, (tuple, list)):
            n = int(size_rng.integers(nodes[0], nodes[1] + 1))
        else:
            n = int(nodes)
        e = c['edges'] if 'edges' in c else c['edges_per_node'] * n
        out.append(make_graph(n, int(e), c['feat'], seed + g, internal=internal))
    return out


def make_batch(cfg, count=None, seed=0, internal=True):
    return Batch.from_data_list(make_graphs(cfg, count, seed, internal))
