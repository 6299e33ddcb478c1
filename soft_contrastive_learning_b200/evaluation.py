"""The callers either side of the retrieval kernels, by the reference's own shapes.

* ``get_top_n``             evaluation/top-n.py:23-119  -- pickles/CSVs in, ``l{l}_dim{d}/<name>.pickle`` out
* ``FeatureCache``          train/train.py:446-454, 1066 -- hard-negative mining: full sort of the feature cache
* ``evaluate_localization`` train/train.py:1181-1185     -- in-training top-5 localization + nearest-by-xy optimum
* ``localization_summary``  train/train.py:360-386       -- "% localized within x m" curves, AUC@Top1, %<rad@Top1

The neighbour searches, the PCA projection, the geographic bookkeeping and the recall curves run on the GPU
(``retrieval.KDTree``, ``netvlad.pca_project``, ``retrieval.geo_topn``, ``retrieval.recall_curves``), and so does the
PCA *fit* (``netvlad.pca_fit``: one exact fit at max(D), every smaller d is a prefix of it; ``pca_solver="sklearn"``
makes the reference's own library call instead); only file IO and list bookkeeping stay on the host.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import formats
from .netvlad import pca_fit, pca_from_sklearn, pca_project
from .retrieval import KDTree, geo_topn, recall_curves, top_n

# top-n.py:34-39: the sweep used for the published checkpoints, and the default cell
FULL_L = (0.0, 0.3, 1.0, 5.0)
FULL_D = (64, 128, 256, 512, 1024, 2048, 4096)


def _cell_pickle(out_root, query_lv_pickle, l, d):
    folder = os.path.join(out_root, "l{}_dim{}".format(l, d))                          # top-n.py:45,82
    name = "".join(os.path.basename(query_lv_pickle).split(".")[:-1])                  # top-n.py:46,84
    return folder, os.path.join(folder, "{}.pickle".format(name))


def get_top_n(pca_lv_pickle, query_lv_pickle, ref_lv_pickle, query_csv, ref_csv, out_root, N=25, L=(0.0,), D=(256,),
              log=print, pca_solver="gpu"):
    """evaluation/top-n.py:23-119 for the sweep ``L`` x ``D`` (defaults: the reference's single cell l=0.0, d=256).

    Skips everything when all cells exist (:41-57) and single cells that exist (:87-89) or have fewer than N
    references left after subsampling (:96-97).  Returns the list of pickles written."""
    if all(os.path.exists(_cell_pickle(out_root, query_lv_pickle, l, d)[1]) for l in L for d in D):
        log("Skipping complete {}".format(query_lv_pickle))
        return []
    if pca_solver not in ("gpu", "sklearn"):
        raise ValueError("pca_solver must be 'gpu' or 'sklearn'")

    full_ref_xy = formats.get_xy(formats.load_csv(ref_csv))                            # :59-62
    full_query_xy = formats.get_xy(formats.load_csv(query_csv))
    pca_f = formats.load_features(pca_lv_pickle)                                       # :65-67
    full_ref_f = formats.load_features(ref_lv_pickle)
    full_query_f = formats.load_features(query_lv_pickle)
    written = []
    if pca_solver == "gpu":                                                            # :74-75, exact, once for the sweep
        v_all, m_all, var_all = pca_fit(pca_f, max(D))
    for d in D:
        log(d)
        if pca_solver == "gpu":
            v, m, var = v_all[:d], m_all, var_all[:d]
        else:
            from sklearn.decomposition import PCA
            v, m, var = pca_from_sklearn(PCA(whiten=True, n_components=d).fit(pca_f))  # the reference's call, :74-75
        with torch.no_grad():                                                          # :76-77 on the tcgen05 GEMM
            pca_ref_f = pca_project(full_ref_f, v, m, var)
            pca_query_f = pca_project(full_query_f, v, m, var)
        for l in L:
            log(l)
            folder, out_pickle = _cell_pickle(out_root, query_lv_pickle, l, d)
            os.makedirs(folder, exist_ok=True)
            if os.path.exists(out_pickle):
                log("{} already exists. Skipping.".format(out_pickle))
                continue
            payload = top_n(pca_ref_f, pca_query_f, full_ref_xy, full_query_xy, N=N, l=l)   # :91-117
            if payload is None:
                continue
            formats.save_pickle(payload, out_pickle)                                   # :119
            written.append(out_pickle)
    return written


class FeatureCache:
    """The mining cache of train/train.py (CACHED_FEATURES / CACHED_FEATURE_INDICES / CACHED_FEATURE_TREE, :1035-1066).

    ``sorted_neighbours(index)`` is lines :446-454: the cached images ordered by feature distance to image ``index``
    (which is first itself), as dataset indices; ``None`` when ``index`` is not in the cache.  The full sort of the
    cache is one exact k = len(cache) query on the GPU index."""

    def __init__(self, features, indices):
        self.features = np.ascontiguousarray(features, dtype=np.float32)
        self.indices = np.asarray(indices)
        if self.features.shape[0] != self.indices.shape[0]:
            raise ValueError("one dataset index per cached feature")
        self.tree = KDTree(self.features)                                              # train.py:1066

    def sorted_neighbours(self, index, k=None):
        hits = np.where(self.indices == index)[0]                                      # :446
        if len(hits) == 0:
            return None
        k = len(self.indices) if k is None else k                                      # MINING_CACHE_SIZE, :449
        sorted_ni = self.tree.query(self.features[hits[0]].reshape(1, -1), k=k, return_distance=False,
                                    sort_results=True)[0]                              # :449-451
        return [self.indices[ni] for ni in sorted_ni]                                  # :452


def evaluate_localization(ref_features, query_features, ref_xy, query_xy, k=5):
    """train/train.py:1181-1185: the k nearest references in feature space and the nearest one on the map.
    Returns (nearest_latent_dists [Q,k], nearest_latent_indices [Q,k], nearest_d_dist [Q,1], nearest_d_indices [Q,1])
    as NumPy arrays like ``KDTree.query``."""
    tree = KDTree(np.asarray(ref_features))
    nearest_latent_dists, nearest_latent_indices = tree.query(np.asarray(query_features), k=k)
    # KDTree(ref_xy).query(query_xy, k=1) in float64: the ground-truth half of the geo kernel
    _, gi, gd = geo_topn(np.asarray(query_xy, dtype=np.float64), np.asarray(ref_xy, dtype=np.float64),
                         nearest_latent_indices[:, :1])
    return (nearest_latent_dists, nearest_latent_indices, gd.cpu().numpy().reshape(-1, 1),
            gi.cpu().numpy().reshape(-1, 1))


def _auc(x, y):
    """sklearn.metrics.auc for increasing x: the trapezoidal rule (train.py:379)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    return float(np.sum((x[1:] - x[:-1]) * (y[1:] + y[:-1]) * 0.5))


def localization_summary(nearest_latent_indices, nearest_d_dist, query_xy, ref_xy, rads=(50, 25, 10), num=25):
    """train/train.py:360-386 without the plotting: for each radius the curves ``Y[n][x]`` = % of queries whose best
    of the first n+1 hits lies within ``X[x]`` metres, the optimum curve from the nearest reference on the map, and the
    two scalars the reference logs (``'{rad}m-auc@Top1'``, ``'%<{rad}m@Top1'``)."""
    tg, _, _ = geo_topn(np.asarray(query_xy, dtype=np.float64), np.asarray(ref_xy, dtype=np.float64),
                        np.asarray(nearest_latent_indices))                            # d_to_nearest_latent, :363-366
    opt = np.asarray(nearest_d_dist, dtype=np.float64).reshape(-1)
    out = {"scalars": {}, "curves": {}}
    for rad in rads:
        X = np.linspace(0, rad, num=num)                                               # :374
        Y = recall_curves(tg, X)                                                       # running min + fraction, :367-376
        Y_opt = np.array([float(np.sum(opt < x)) / float(len(opt)) * 100 for x in X])  # :386
        out["scalars"]["{}m-auc@Top1".format(rad)] = _auc(X, Y[0])                     # :379-380
        out["scalars"]["%<{}m@Top1".format(rad)] = float(Y[0][-1])                     # :383-384
        out["curves"][rad] = {"X": X, "Y": Y, "optimum": Y_opt}
    return out
