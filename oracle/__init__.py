"""CPU oracle for the DeepRank-GNN hot path.  TEST INFRASTRUCTURE ONLY.

This package is a pure-torch / numpy float32 CPU *restatement* of the reference
algorithm (DeepRank/Deeprank-GNN v0.1.4) for the GNN forward/backward hot path:

* ``pyg_min``   - the un-vendored third-party primitives the reference reaches
                  (torch_scatter ``scatter_sum/mean/max``, torch_geometric
                  ``Batch.from_data_list``, ``consecutive_cluster``, ``pool_edge``,
                  ``pool_batch``, ``max_pool_x``, ``inits.uniform``; torch_sparse
                  ``coalesce``), restated from their published behaviour for the
                  torch-1.8-era versions pinned in the reference CI
                  (``.github/workflows/build.yml:36-48``: torch==1.8.0 + matching
                  torch-scatter 2.0.x / torch-sparse 0.6.x / torch-geometric 1.7.x).
* ``pooling``   - ``deeprank_gnn/community_pooling.py:25-30, 161-251``.
* ``nets``      - ``deeprank_gnn/ginet.py:22-141``, ``sGAT.py:19-138``,
                  ``foutnet.py:15-125``.
* ``step``      - ``deeprank_gnn/NeuralNet.py:239-263, 477-503, 616-631``.
* ``mcl``       - ``markov_clustering`` (un-vendored; ``community_pooling.py:142-155``).

PARITY PINNING.  The reference cannot be imported in this environment
(torch_geometric / torch_scatter / torch_sparse / h5py / markov_clustering are
not installed and there is no network) and its own tests assert no values
(every test is a smoke test).  The oracle is therefore pinned only by
  (1) the hand-computable toy graph of ``tests/test_community_pooling.py:10-18``,
  (2) structural facts of the shipped fixture ``tests/hdf5/1ATN_residue.hdf5``
      (node / cluster counts, ``len(depth_1) == n_unique(depth_0)``) and the 20
      MCL cluster vectors stored in it (they pin ``mcl``),
  (3) the state_dict layout of the 13 shipped checkpoints,
  (4) algebraic identities of the reference code (alpha == 1, closed forms).
For the floating-point network outputs: **parity unpinned** at the third-party
boundary (stated in DESIGN.md as well).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import this package.  The product package
(``deeprank-gnn_b200``) never does.
"""
