"""soft_contrastive_learning_b200 -- the descriptor-space hot path of janinethoma/soft_contrastive_learning on B200.

  losses     wms_loss, ms_loss, triplet/quadruplet family, Huber-distance triplet, logratio, get_loss(name)
  netvlad    netVLAD(inputs, assignment_kernel, cluster_centers), pca_project(x, v, m, var)
  retrieval  KDTree(ref).query(q, k), ShardedKDTree, top_n(...), recall_at_n(...)
  evaluation get_top_n(pickles, csvs -> top-N pickles), FeatureCache (mining), evaluate_localization, localization_summary
  formats    load/save_pickle, load/save_csv, get_xy, load/save_features (util/io.py, util/meta.py formats)

All arithmetic runs in libscl_b200.so (hand-written sm_100a CUDA behind the C ABI of include/scl_b200.h).
Importing this package does not load the library; the first call does, and raises if it is missing.
"""
__version__ = "0.1.0"

from . import synth  # noqa: F401  (NumPy-only workload generators)


def __getattr__(name):
    import importlib
    if name in ("losses", "netvlad", "retrieval", "evaluation", "formats", "_lib", "build"):
        return importlib.import_module(f"{__name__}.{name}")
    raise AttributeError(name)
