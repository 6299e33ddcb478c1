"""CPU oracle for the descriptor-space hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, on the CPU (torch-CPU float64 / NumPy / scikit-learn), the
arithmetic of the reference's hot path so the CUDA product can be checked against it:

  oracle.losses     <- /root/reference/model/losses.py (+ the un-vendored pointnetvlad_cls losses)
  oracle.netvlad    <- netvlad_tf.layers.netVLAD as called at model/nets.py:66-67, PCA op train/train.py:646-652
  oracle.retrieval  <- evaluation/top-n.py:69-119, evaluation/roc.py:200-216, train/train.py:363-386

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  The product package
``soft_contrastive_learning_b200`` never imports it and has no CPU fallback.

Parity pin: the reference ships no tests or golden vectors.  The
oracle is pinned instead by executing the reference's *own source file*
``model/losses.py`` over a NumPy shim of the handful of TF-1.x ops it uses
(``tests/golden/tf_numpy_shim.py``, ``tests/golden/make_golden.py``) and freezing the
outputs under ``tests/golden/``; retrieval is pinned against the reference's actual
library call ``sklearn.neighbors.KDTree.query``.  The NetVLAD layer and the four
pointnetvlad losses live in un-vendored, un-pinned third-party repos (README.md:8-12)
and are restated from their published algorithm: parity for those two is UNPINNED
beyond the structural twins that exist in-repo (``evil_triplet_loss`` etc.).
"""
