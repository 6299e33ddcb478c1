"""Oracle (TEST INFRASTRUCTURE): brute-force top-N retrieval, geo bookkeeping and recall@N.

R1  /root/reference/evaluation/top-n.py:103-108 -- ``KDTree(ref).query(query, k=N, return_distance=True,
    sort_results=True)``: exact Euclidean kNN, float64 inside, ascending.  ``knn_kdtree`` runs that very
    library call (scikit-learn is installed here and on the GPU box); ``knn_bruteforce`` is the float64
    restatement with the explicit ``(dist, idx)`` tie-break the CUDA path must reproduce bit-exactly.
R2  top-n.py:69, 91-100, 110-119 -- xy distances, greedy reference subsampling, top_g_dists, gt.
R3  evaluation/roc.py:200-216 and train/train.py:363-386 -- recall curves.
"""
from __future__ import annotations

import numpy as np


def knn_kdtree(ref_f, query_f, k):
    """The reference's own call (top-n.py:103-106)."""
    from sklearn.neighbors import KDTree
    tree = KDTree(ref_f)
    top_f_dists, top_i = tree.query(query_f, k=k, return_distance=True, sort_results=True)
    return np.asarray(top_f_dists), np.asarray(top_i, dtype=np.int64)


def knn_bruteforce(ref_f, query_f, k, chunk=256):
    """float64 direct-difference distances (what KDTree's leaf scan computes), sorted by (dist, idx)."""
    ref = np.asarray(ref_f, dtype=np.float64)
    qry = np.asarray(query_f, dtype=np.float64)
    Q, R = qry.shape[0], ref.shape[0]
    k = min(k, R)
    out_d = np.empty((Q, k), dtype=np.float64)
    out_i = np.empty((Q, k), dtype=np.int64)
    rn = (ref * ref).sum(1)
    for q0 in range(0, Q, chunk):
        qq = qry[q0:q0 + chunk]
        # Gram form only to shortlist; the shortlist is then rescored by direct differences
        approx = rn[None, :] - 2.0 * (qq @ ref.T)
        kk = min(R, max(4 * k, k + 64))
        cand = np.argpartition(approx, kk - 1, axis=1)[:, :kk]
        for j in range(qq.shape[0]):
            c = cand[j]
            d2 = ((ref[c] - qq[j]) ** 2).sum(1)
            # guard: anything outside the shortlist must be farther than the k-th exact (slack for Gram rounding)
            order = np.lexsort((c, d2))[:k]
            out_d[q0 + j] = np.sqrt(d2[order])
            out_i[q0 + j] = c[order]
    return out_d, out_i


def knn_sgemm_allcores(ref_f, query_f, k):
    """The strongest plain-CPU form of top-n.py:103-106 (bench baseline only): float32 ``|r|^2 - 2 q.r`` through the
    multi-threaded BLAS sgemm for a shortlist of 4k candidates, float64 direct-difference rescore, (dist, idx) order."""
    ref = np.ascontiguousarray(ref_f, dtype=np.float32)
    qry = np.ascontiguousarray(query_f, dtype=np.float32)
    Q, R = qry.shape[0], ref.shape[0]
    k = min(k, R)
    kk = min(R, max(4 * k, k + 64))
    approx = (ref * ref).sum(1)[None, :] - 2.0 * (qry @ ref.T)
    cand = np.argpartition(approx, kk - 1, axis=1)[:, :kk]
    out_d = np.empty((Q, k), dtype=np.float64)
    out_i = np.empty((Q, k), dtype=np.int64)
    for j in range(Q):
        c = cand[j]
        d2 = ((ref[c].astype(np.float64) - qry[j].astype(np.float64)) ** 2).sum(1)
        order = np.lexsort((c, d2))[:k]
        out_d[j], out_i[j] = np.sqrt(d2[order]), c[order]
    return out_d, out_i


def knn_bruteforce_exact(ref_f, query_f, k):
    """Tiny-problem version with no shortlist at all: full float64 direct differences."""
    ref = np.asarray(ref_f, dtype=np.float64)
    qry = np.asarray(query_f, dtype=np.float64)
    k = min(k, ref.shape[0])
    out_d = np.empty((qry.shape[0], k))
    out_i = np.empty((qry.shape[0], k), dtype=np.int64)
    idx = np.arange(ref.shape[0])
    for j in range(qry.shape[0]):
        d2 = ((ref - qry[j]) ** 2).sum(1)
        order = np.lexsort((idx, d2))[:k]
        out_d[j] = np.sqrt(d2[order])
        out_i[j] = order
    return out_d, out_i


def subsample_refs(ref_xy, l):
    """top-n.py:91-94: greedy keep of refs at least ``l`` metres from the last kept one."""
    ref_idx = [0]
    for i in range(len(ref_xy)):
        if sum((ref_xy[i, :] - ref_xy[ref_idx[-1], :]) ** 2) >= l ** 2:
            ref_idx.append(i)
    return ref_idx


def top_n(ref_f, query_f, ref_xy, query_xy, N=25, l=0.0, knn=knn_kdtree):
    """top-n.py:69,91-119 -> [top_i, top_g_dists, top_f_dists, gt_i, gt_g_dist, ref_idx] (the pickle payload)."""
    from sklearn.metrics import pairwise_distances
    full_xy_dists = pairwise_distances(query_xy, ref_xy, metric="euclidean")         # :69
    ref_idx = subsample_refs(ref_xy, l)                                               # :91-94
    if len(ref_idx) < N:                                                              # :96-97
        return None
    sub_f = np.array([ref_f[i, :] for i in ref_idx])                                  # :99
    xy_dists = np.array([full_xy_dists[:, i] for i in ref_idx]).transpose()           # :100
    top_f_dists, top_i = knn(sub_f, query_f, N)                                       # :103-108
    num_q = query_xy.shape[0]
    top_g_dists = [[xy_dists[q, r] for r in top_i[q, :]] for q in range(num_q)]       # :110
    gt_i = np.argmin(xy_dists, axis=1)                                                # :112
    gt_g_dist = np.min(xy_dists, axis=1)                                              # :113
    top_i = [[ref_idx[r] for r in top_i[q, :]] for q in range(num_q)]                 # :116
    gt_i = [ref_idx[r] for r in gt_i]                                                 # :117
    return [top_i, top_g_dists, top_f_dists, gt_i, gt_g_dist, ref_idx]


def recall_curve_top1(top_g_dists, t=25.0, num=50):
    """roc.py:213-216: % of queries whose top-1 hit lies within x metres, x in linspace(0,t,num)."""
    t_1_d = np.array([td[0] for td in top_g_dists])
    X = np.linspace(0, t, num=num)
    Y = [float(sum(t_1_d < x)) / float(len(t_1_d)) * 100 for x in X]
    return X, np.array(Y)


def recall_upper_bound(gt_g_dist, t=25.0, num=50):
    """roc.py:200-201."""
    gt_g_dist = np.asarray(gt_g_dist)
    X = np.linspace(0, t, num=num)
    Y = [float(sum(gt_g_dist < x)) / float(len(gt_g_dist)) * 100 for x in X]
    return X, np.array(Y)


def recall_at_n(top_g_dists, rad=25.0, num=25):
    """train/train.py:368-375: top_n[q,j] = min(d[q,0..j]); Y_n(x) = % of queries with top_n[:,n] < x."""
    d = np.asarray(top_g_dists, dtype=np.float64)
    top_n_ = np.minimum.accumulate(d, axis=1)
    X = np.linspace(0, rad, num=num)
    Y = np.array([[float(sum(top_n_[:, n] < x)) / float(len(top_n_[:, n])) * 100 for x in X]
                  for n in range(top_n_.shape[1])])
    return X, Y


def localization_summary(nearest_latent_indices, nearest_d_dist, query_xy, ref_xy, rads=(50, 25, 10)):
    """train/train.py:360-386 (evaluate_localization_thread) without the plotting, loops as written there."""
    nq = nearest_latent_indices.shape[0]
    d_to_nearest_latent = np.empty(nearest_latent_indices.shape)
    for i in range(nq):                                                               # :364-367
        for j in range(nearest_latent_indices.shape[1]):
            other_index = nearest_latent_indices[i][j]
            d_to_nearest_latent[i, j] = np.linalg.norm(query_xy[i, :] - ref_xy[other_index, :])
    top_n_ = np.empty(nearest_latent_indices.shape)
    for i in range(nq):                                                               # :369-371
        for j in range(nearest_latent_indices.shape[1]):
            top_n_[i, j] = min(d_to_nearest_latent[i, 0:(j + 1)])
    from sklearn.metrics import auc
    out = {"scalars": {}, "curves": {}}
    for rad in rads:                                                                  # :372-386
        X = np.linspace(0, rad, num=25)
        Ys = []
        for n in range(top_n_.shape[1]):
            Y = [float(sum(top_n_[:, n] < x)) / float(len(top_n_[:, n])) * 100 for x in X]
            Ys.append(Y)
            if n == 0:
                out["scalars"]["{}m-auc@Top1".format(rad)] = auc(X, Y)
                out["scalars"]["%<{}m@Top1".format(rad)] = Y[-1]
        Yo = [float(sum(np.array(nearest_d_dist).reshape(-1) < x)) / float(len(top_n_[:, 0])) * 100 for x in X]
        out["curves"][rad] = {"X": X, "Y": np.array(Ys), "optimum": np.array(Yo)}
    return out


def mining_sorted_neighbours(cached_features, cached_indices, index, k):
    """train/train.py:446-452: cache entries ordered by feature distance to image ``index`` (sklearn KDTree)."""
    from sklearn.neighbors import KDTree
    fis = np.where(cached_indices == index)[0]
    if len(fis) == 0:
        return None
    tree = KDTree(cached_features)                                                    # :1066
    sorted_ni = tree.query(cached_features[fis[0], :].reshape(1, -1), k=k, return_distance=False, sort_results=True)[0]
    return [cached_indices[ni] for ni in sorted_ni]
