"""Markov clustering as the reference runs it at pre-clustering time.  TEST INFRASTRUCTURE ONLY.

``deeprank_gnn/community_pooling.py:95-158`` (``community_detection``, method ``'mcl'``) builds an undirected
networkx graph from ``internal_edge_index`` (no weights: ``DataSet.py:76,84`` pass no ``edge_attr``), converts it
with ``nx.to_scipy_sparse_matrix`` and calls the third-party package ``markov_clustering`` (un-vendored, not
installed here; the reference pins no version - the API used, ``mc.run_mcl(matrix)`` + ``mc.get_clusters``, is
that of markov_clustering 0.0.6, the only release line with these names).  This file restates that package's
published algorithm for the DEFAULT parameters the reference uses:

    run_mcl(matrix, expansion=2, inflation=2, loop_value=1, iterations=100, pruning_threshold=0.001,
            pruning_frequency=1, convergence_check_frequency=1)
      M <- A with M[i,i] = loop_value            (add_self_loops)
      M <- column-normalise(M)                   (sklearn normalize(norm='l1', axis=0))
      repeat <= 100 times:
          last <- M
          M <- M @ M                             (expand, power 2)
          M <- column-normalise(M ** 2)          (inflate, element-wise power 2)
          M <- prune(M, 0.001)                   (drop entries < threshold, but keep every column's maximum)
          stop when allclose(M, last, rtol=1e-5, atol=1e-8)
    get_clusters(M): attractors = rows with a non-zero diagonal; cluster of an attractor = the columns with a
      non-zero entry in its row; clusters = sorted(set(tuples))

and the labelling of ``community_detection``: ``index[list(c)] = ic`` in that sorted order (a node that sits in
several clusters keeps the LAST one; ids may therefore have gaps).

PINNED on the 20 cluster vectors stored in the shipped fixture (``clustering/mcl/depth_{0,1}`` of the 10 graphs of
``tests/hdf5/1ATN_residue.hdf5``, written by the reference's ``PreCluster``): ``tests/test_mcl.py``.
"""
import numpy as np
import scipy.sparse as sp


def adjacency(edge_index, num_nodes):
    """``nx.Graph`` + ``nx.to_scipy_sparse_matrix``: undirected, one unit-weight entry per distinct pair (both
    triangles), duplicate / reversed edges collapse, a self-loop gives one diagonal entry."""
    ei = np.asarray(edge_index, dtype=np.int64).reshape(2, -1)
    a = sp.lil_matrix((num_nodes, num_nodes), dtype=np.float64)
    for i, j in zip(ei[0].tolist(), ei[1].tolist()):
        a[i, j] = 1.0
        a[j, i] = 1.0
    return a.tocsc()


def _normalize(m):
    """sklearn.preprocessing.normalize(m, norm='l1', axis=0) on a CSC matrix: divide every column by the sum of
    its absolute values (columns that sum to 0 are left alone)."""
    m = sp.csc_matrix(m, dtype=np.float64, copy=True)
    sums = np.asarray(abs(m).sum(axis=0)).ravel()
    sums[sums == 0.0] = 1.0
    m.data /= np.repeat(sums, np.diff(m.indptr))
    return m


def _prune(m, threshold):
    pruned = sp.dok_matrix(m.shape)
    keep = m >= threshold
    pruned[keep] = m[keep]
    pruned = pruned.tocsc()
    num_cols = m.shape[1]
    row_indices = np.asarray(m.argmax(axis=0)).reshape((num_cols,))
    col_indices = np.arange(num_cols)
    pruned[row_indices, col_indices] = m[row_indices, col_indices]
    return sp.csc_matrix(pruned)


def _allclose(a, b, rtol=1e-5, atol=1e-8):
    c = np.abs(a - b) - rtol * np.abs(b)
    return c.max() <= atol


def run_mcl(matrix, expansion=2, inflation=2, loop_value=1, iterations=100, pruning_threshold=0.001):
    m = sp.lil_matrix(matrix, dtype=np.float64)
    if loop_value > 0:
        m.setdiag(float(loop_value))
    m = _normalize(m.tocsc())
    for _ in range(iterations):
        last = m.copy()
        m = m ** expansion
        m = _normalize(m.power(inflation))
        if pruning_threshold > 0:
            m = _prune(m, pruning_threshold)
        if _allclose(m, last):
            break
    return m


def get_clusters(matrix):
    m = sp.csc_matrix(matrix)
    attractors = m.diagonal().nonzero()[0]
    clusters = set()
    for attractor in attractors:
        clusters.add(tuple(m.getrow(attractor).nonzero()[1].tolist()))
    return sorted(list(clusters))


def community_detection_mcl(edge_index, num_nodes):
    """``community_detection(edge_index, num_nodes, method='mcl')`` (community_pooling.py:142-155) -> int64 [num_nodes]."""
    clusters = get_clusters(run_mcl(adjacency(edge_index, num_nodes)))
    index = np.zeros(num_nodes, dtype=np.int64)
    for ic, c in enumerate(clusters):
        index[list(c)] = ic
    return index


def run_mcl_dense(a, iterations=100, threshold=0.001):
    """The same iteration on a dense float64 matrix with explicit loops in the summation order of the sparse
    product (column j of M @ M accumulates M[:, k] * M[k, j] over the non-zero k in ascending order) - the form
    the CUDA kernel implements; returns (M, iterations run)."""
    m = np.array(a, dtype=np.float64)
    n = m.shape[0]
    np.fill_diagonal(m, 1.0)
    s = np.abs(m).sum(axis=0)
    s[s == 0] = 1.0
    m = m / s
    it = 0
    for it in range(1, iterations + 1):
        last = m
        e = np.zeros_like(m)
        for j in range(n):
            ks = np.nonzero(last[:, j])[0]
            acc = np.zeros(n)
            for k in ks:
                acc = acc + last[:, k] * last[k, j]
            e[:, j] = acc
        e = e * e
        s = np.abs(e).sum(axis=0)
        s[s == 0] = 1.0
        e = e / s
        am = e.argmax(axis=0)
        p = np.where(e >= threshold, e, 0.0)
        p[am, np.arange(n)] = e[am, np.arange(n)]
        m = p
        if (np.abs(m - last) - 1e-5 * np.abs(last)).max() <= 1e-8:
            break
    return m, it
