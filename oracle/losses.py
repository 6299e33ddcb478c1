"""Oracle (TEST INFRASTRUCTURE): torch-CPU restatement of the reference's metric-learning losses.

Every function follows the cited lines of /root/reference/model/losses.py op by op, with the
TF-1.x op replaced by the torch op of identical value *and* gradient convention:

  tf.maximum(x, 0)        -> torch.clamp(x, min=0)   (gradient passes when x >= 0, ties included)
  tf.reduce_max/min       -> torch.amax/amin         (gradient split evenly among ties)
  tf.where(c, a, b)       -> torch.where             (no gradient through c)
  tf.nn.l2_normalize(x,1) -> x * rsqrt(max(sum x^2, 1e-12))
  tf.losses.huber_loss    -> delta=1, mean over all elements (SUM_BY_NONZERO_WEIGHTS, unit weights)

Gradients come from torch.autograd on float64; tests check them against central finite
differences of the reference source itself run over the NumPy TF shim.

The four ``pointnetvlad_cls`` losses (triplet / lazy_triplet / quadruplet / lazy_quadruplet)
are NOT under /root/reference (README.md:11, un-pinned); they are restated from the published
PointNetVLAD algorithm, whose structure the in-repo twins ``evil_triplet_loss``
(losses.py:63-73) and ``evil_quadruplet_loss`` (losses.py:197-214) share exactly except for
reduce_max in place of reduce_min over positives (losses.py:217-222).
"""
from __future__ import annotations

import numpy as np
import torch

F64 = torch.float64


def _t(x, dtype=F64):
    if isinstance(x, torch.Tensor):
        return x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def l2_normalize(x, dim):
    """tf.nn.l2_normalize: x * rsqrt(max(sum(x^2), 1e-12)) (losses.py:7, :84; nets.py:66)."""
    ss = (x * x).sum(dim=dim, keepdim=True)
    return x * torch.rsqrt(torch.clamp(ss, min=1e-12))


# --------------------------------------------------------------------------------------
# W1: wms_loss  (model/losses.py:5-60)
# --------------------------------------------------------------------------------------
def wms_masks(distances, d_alpha, d_beta, wfunction="exp"):
    """Soft positive / negative weights from GPS distances (losses.py:11-19), evaluated in the dtype of
    ``distances``.  ``wms_loss`` calls this in FLOAT32, like the reference (see there)."""
    if wfunction == "lin":      # losses.py:11-13
        mask_pos = torch.where(distances < d_beta, 1.0 - distances / d_beta, torch.zeros_like(distances))
        mask_neg = torch.where(distances < d_beta, distances / d_beta, torch.ones_like(distances))
    elif wfunction == "tanh":   # losses.py:14-16
        # float32 tanh saturates to exactly 1.0 near x ~ 9 and faithful implementations disagree on where (any x > 8.66
        # may round either way), which flips `mask_pos > 0`.  The oracle pins the correctly rounded value: tanh in
        # float64, rounded once to the mask dtype.  (TF-1.10's Eigen tanh may saturate elsewhere: unpinnable here.)
        t = torch.tanh((distances / d_beta).to(torch.float64)).to(distances.dtype)
        mask_pos = 1.0 - t
        mask_neg = t
    else:                       # 'exp' default, losses.py:17-19
        mask_pos = 1.0 / (1.0 + torch.exp(d_alpha * (distances - d_beta)))
        mask_neg = 1.0 / (1.0 + torch.exp(d_alpha * (d_beta - distances)))
    return mask_pos, mask_neg


def _ms_core(sim_mat, mask_pos, mask_neg, alpha, beta, lamb, eps, ms_mining, sumfunction, return_masks=False):
    """Shared body of wms_loss (losses.py:26-58) and ms_loss (losses.py:95-120)."""
    sim_mat = torch.clamp(sim_mat, min=0.0)                                   # :26 / :95
    pos_mat = sim_mat * mask_pos                                              # :28
    neg_mat = sim_mat * mask_neg                                              # :29
    if ms_mining:                                                             # :31-37
        max_val = torch.amax(neg_mat, dim=1, keepdim=True)
        tmp_max_val = torch.amax(pos_mat, dim=1, keepdim=True)
        min_val = torch.amin((sim_mat - tmp_max_val) * mask_pos, dim=1, keepdim=True) + tmp_max_val
        # comparison outputs carry no gradient in TF: thresholds are gradient-free
        mask_pos = torch.where((pos_mat < max_val + eps).detach(), mask_pos, torch.zeros_like(mask_pos))
        mask_neg = torch.where((neg_mat > min_val - eps).detach(), mask_neg, torch.zeros_like(mask_neg))
    if sumfunction == "plain":                                                # :39-46
        pos_exp = torch.where(mask_pos > 0.0, pos_mat, torch.zeros_like(pos_mat))
        neg_exp = torch.where(mask_neg > 0.0, neg_mat, torch.zeros_like(neg_mat))
        loss = (neg_exp.sum(dim=1) - pos_exp.sum(dim=1)).mean()
    else:                                                                     # 'ms', :48-58
        pos_exp = torch.exp(-alpha * (pos_mat - lamb))
        pos_exp = torch.where(mask_pos > 0.0, pos_exp, torch.zeros_like(pos_exp))
        neg_exp = torch.exp(beta * (neg_mat - lamb))
        neg_exp = torch.where(mask_neg > 0.0, neg_exp, torch.zeros_like(neg_exp))
        pos_term = torch.log(1.0 + pos_exp.sum(dim=1)) / alpha
        neg_term = torch.log(1.0 + neg_exp.sum(dim=1)) / beta
        loss = (pos_term + neg_term).mean()
    if return_masks:
        return loss, (mask_pos > 0.0), (mask_neg > 0.0)
    return loss


def wms_loss(distances, embeddings, d_alpha, d_beta, alpha=2.0, beta=50.0, lamb=1.0, eps=0.1,
             ms_mining=True, wfunction="exp", sumfunction="ms", return_masks=False):
    """The paper's weighted multi-similarity loss, flat form: distances [B,B], embeddings [B,D].

    model/losses.py:5-60.  Call site train/train.py:852 passes d_alpha=ALPHA, d_beta=BETA only.
    """
    embeddings = _keep(embeddings)
    embeddings = l2_normalize(embeddings, 1)                                  # :7
    batch_size = embeddings.shape[0]                                          # :9
    # The masks are FLOAT32 in the reference: the distances placeholder is float32 (train.py:684-686) and :22
    # casts to float32 before subtracting the identity.  This is not a rounding detail: tf.exp overflows to inf in
    # float32 beyond ~126 m, which makes mask_pos exactly 0 there, and the later `mask_pos > 0` tests (:50) then
    # drop those pairs from the positive sum -- in float64 the same pairs would each contribute e^{alpha*lamb}.
    d32 = _t(distances, torch.float32)
    mask_pos, mask_neg = wms_masks(d32, d_alpha, d_beta, wfunction)           # :11-19, float32
    mask_pos = mask_pos - torch.eye(batch_size, dtype=torch.float32)          # :22, float32
    mask_pos, mask_neg = mask_pos.to(embeddings.dtype), mask_neg.to(embeddings.dtype)
    sim_mat = embeddings @ embeddings.T                                       # :25
    return _ms_core(sim_mat, mask_pos, mask_neg, alpha, beta, lamb, eps, ms_mining, sumfunction, return_masks)


def wms_loss_tuples(distances, embeddings, d_alpha, d_beta, **kw):
    """Tuple-mode batching: distances [T,S,S], embeddings [T,S,D] -> mean over
    tuples of the per-tuple ``wms_loss``.  For T=1 this is exactly what train.py:684-686,852 feeds
    (the reference's [1,S,S] x [S,D] broadcast gives the same value as the 2-D call)."""
    T = embeddings.shape[0]
    losses = [wms_loss(distances[t], embeddings[t], d_alpha, d_beta, **kw) for t in range(T)]
    return torch.stack(losses).mean()


# --------------------------------------------------------------------------------------
# W2: ms_loss  (model/losses.py:76-122); labels built at train/train.py:821-826
# --------------------------------------------------------------------------------------
def ms_labels(T, P, N):
    """train/train.py:822-826: anchor+positives share one class per tuple, each negative its own."""
    one = np.concatenate((np.zeros(1 + P), np.arange(N) + 1))
    allc = one
    for b in range(1, T):
        allc = np.concatenate((allc, one + b * (N + 1)))
    return allc


def ms_loss(labels, embeddings, alpha=2.0, beta=50.0, lamb=1.0, eps=0.1, ms_mining=True, return_masks=False):
    """Multi-similarity loss with hard label masks, model/losses.py:76-122."""
    embeddings = _keep(embeddings)
    embeddings = l2_normalize(embeddings, 1)                                  # :84
    labels = torch.as_tensor(np.asarray(labels)).reshape(-1, 1)               # :85
    batch_size = embeddings.shape[0]
    adjacency = labels == labels.T                                            # :89
    mask_pos = adjacency.to(embeddings.dtype) - torch.eye(batch_size, dtype=embeddings.dtype)  # :92
    mask_neg = (~adjacency).to(embeddings.dtype)                              # :93
    sim_mat = embeddings @ embeddings.T                                       # :95
    return _ms_core(sim_mat, mask_pos, mask_neg, alpha, beta, lamb, eps, ms_mining, "ms", return_masks)


# --------------------------------------------------------------------------------------
# L1-L4: pointnetvlad_cls family (external) + in-repo evil_* twins (losses.py:63-73,197-222)
# --------------------------------------------------------------------------------------
def _sqdist(a, b):
    """sum(squared_difference(b, tile(a)), 2): a [T,1,D], b [T,n,D] -> [T,n]."""
    return ((b - a) ** 2).sum(dim=2)


def best_pos_distance(query, pos_vecs):
    """pointnetvlad_cls.best_pos_distance: reduce_min over positives."""
    return torch.amin(_sqdist(query, pos_vecs), dim=1)


def worst_pos_distance(query, pos_vecs):
    """model/losses.py:217-222: reduce_max over positives."""
    return torch.amax(_sqdist(query, pos_vecs), dim=1)


def _hinge_terms(ref_pos, a, neg_vecs, margin):
    # tf.maximum(m + best_pos - ||neg - a||^2, 0)   (losses.py:70-72)
    return torch.clamp(margin + ref_pos[:, None] - _sqdist(a, neg_vecs), min=0.0)


def triplet_loss(q_vec, pos_vecs, neg_vecs, margin):
    """pointnetvlad_cls.triplet_loss (call train/train.py:701): mean_t sum_n hinge."""
    q_vec, pos_vecs, neg_vecs = map(_keep, (q_vec, pos_vecs, neg_vecs))
    return _hinge_terms(best_pos_distance(q_vec, pos_vecs), q_vec, neg_vecs, margin).sum(dim=1).mean()


def lazy_triplet_loss(q_vec, pos_vecs, neg_vecs, margin):
    """pointnetvlad_cls.lazy_triplet_loss (call train/train.py:703): mean_t max_n hinge."""
    q_vec, pos_vecs, neg_vecs = map(_keep, (q_vec, pos_vecs, neg_vecs))
    return torch.amax(_hinge_terms(best_pos_distance(q_vec, pos_vecs), q_vec, neg_vecs, margin), dim=1).mean()


def quadruplet_loss(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2):
    """pointnetvlad_cls.quadruplet_loss (call train/train.py:707-708)."""
    q_vec, pos_vecs, neg_vecs, other_neg = map(_keep, (q_vec, pos_vecs, neg_vecs, other_neg))
    best = best_pos_distance(q_vec, pos_vecs)
    second = _hinge_terms(best, other_neg, neg_vecs, m2).sum(dim=1).mean()
    return triplet_loss(q_vec, pos_vecs, neg_vecs, m1) + second


def lazy_quadruplet_loss(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2):
    """pointnetvlad_cls.lazy_quadruplet_loss (call train/train.py:710-712)."""
    q_vec, pos_vecs, neg_vecs, other_neg = map(_keep, (q_vec, pos_vecs, neg_vecs, other_neg))
    best = best_pos_distance(q_vec, pos_vecs)
    second = torch.amax(_hinge_terms(best, other_neg, neg_vecs, m2), dim=1).mean()
    return lazy_triplet_loss(q_vec, pos_vecs, neg_vecs, m1) + second


def evil_triplet_loss(q_vec, pos_vecs, neg_vecs, margin):
    """model/losses.py:63-73 (worst positive instead of best)."""
    q_vec, pos_vecs, neg_vecs = map(_keep, (q_vec, pos_vecs, neg_vecs))
    return _hinge_terms(worst_pos_distance(q_vec, pos_vecs), q_vec, neg_vecs, margin).sum(dim=1).mean()


def evil_quadruplet_loss(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2):
    """model/losses.py:197-214."""
    q_vec, pos_vecs, neg_vecs, other_neg = map(_keep, (q_vec, pos_vecs, neg_vecs, other_neg))
    worst = worst_pos_distance(q_vec, pos_vecs)
    second = _hinge_terms(worst, other_neg, neg_vecs, m2).sum(dim=1).mean()
    return evil_triplet_loss(q_vec, pos_vecs, neg_vecs, m1) + second


def _keep(x):
    return x if isinstance(x, torch.Tensor) else _t(x)


# --------------------------------------------------------------------------------------
# L5: distance_triplet_loss with Huber distance term (losses.py:233-264, 678-690)
# --------------------------------------------------------------------------------------
def scale_distances(a_feature, pos_feature, squared_d_dists, d_max_squared, f_max_squared):
    """model/losses.py:678-690."""
    squared_f_dists = _sqdist(a_feature, pos_feature)
    return squared_d_dists / d_max_squared, squared_f_dists / f_max_squared


def huber(labels, predictions, delta=1.0, reduce=True):
    """tf.losses.huber_loss(labels, predictions, delta=1.0): 0.5*q^2 + delta*(|e|-q), q=min(|e|,delta)."""
    err = predictions - labels
    abs_err = err.abs()
    quad = torch.clamp(abs_err, max=delta)
    lin = abs_err - quad
    losses = 0.5 * quad * quad + delta * lin
    return losses.mean() if reduce else losses


def distance_loss(a_feature, pos_feature, squared_d_dists, d_max_squared, f_max_squared):
    """model/losses.py:225-230."""
    a_feature, pos_feature, squared_d_dists = map(_keep, (a_feature, pos_feature, squared_d_dists))
    sd, sf = scale_distances(a_feature, pos_feature, squared_d_dists, d_max_squared, f_max_squared)
    return ((sf - sd) ** 2).mean(dim=1).mean(dim=0)


def huber_distance_loss(a_feature, pos_feature, squared_d_dists, d_max_squared, f_max_squared):
    """model/losses.py:233-236: huber_loss(labels=scaled_d, predictions=scaled_f)."""
    a_feature, pos_feature, squared_d_dists = map(_keep, (a_feature, pos_feature, squared_d_dists))
    sd, sf = scale_distances(a_feature, pos_feature, squared_d_dists, d_max_squared, f_max_squared)
    return huber(sd, sf)


_TRIPLETS = {"triplet_loss": triplet_loss, "lazy_triplet_loss": lazy_triplet_loss}


def distance_triplet_loss(a_feature, pos_features, neg_features, margin, lam, squared_d_dists, d_max_squared,
                          f_max_squared, triplet_loss_name="triplet_loss", distance_loss_name="huber_distance_loss"):
    """model/losses.py:239-264; getattr(pointnetvlad_cls, triplet_loss_name) dispatch at :256,:261."""
    trip = _TRIPLETS[triplet_loss_name](a_feature, pos_features, neg_features, margin)
    if "huber" in distance_loss_name:
        return trip + lam * huber_distance_loss(a_feature, pos_features, squared_d_dists, d_max_squared, f_max_squared)
    return trip + lam * distance_loss(a_feature, pos_features, squared_d_dists, d_max_squared, f_max_squared)


def _best_distance(a_feature, pos_features, squared_d_dists, d_max_squared, f_max_squared):
    """model/losses.py:664-668."""
    sd, sf = scale_distances(a_feature, pos_features, squared_d_dists, d_max_squared, f_max_squared)
    return ((sf - sd) ** 2).min(dim=1).values


def _best_huber_distance(a_feature, pos_features, squared_d_dists, d_max_squared, f_max_squared):
    """model/losses.py:671-675: huber_loss(labels=scaled_f, predictions=scaled_d, reduction=NONE), min over positives."""
    sd, sf = scale_distances(a_feature, pos_features, squared_d_dists, d_max_squared, f_max_squared)
    return huber(sf, sd, reduce=False).min(dim=1).values


def distance_quadruplet_loss(a_feature, pos_features, neg_features, other_neg, m1, m2, lam, squared_d_dists,
                             d_max_squared, f_max_squared, triplet_loss_name="triplet_loss",
                             distance_loss_name="huber_distance_loss"):
    """model/losses.py:267-307 (calls train/train.py:729-763): distance_triplet_loss + a second hinge in which the
    "best positive" is the smallest distance-term element and the negative<->other distances are scaled by f_max;
    the second hinge always takes the max over negatives (:301-305), whatever the triplet name."""
    a_feature, pos_features, neg_features, other_neg, squared_d_dists = map(
        _keep, (a_feature, pos_features, neg_features, other_neg, squared_d_dists))
    trip = distance_triplet_loss(a_feature, pos_features, neg_features, m1, lam, squared_d_dists, d_max_squared,
                                 f_max_squared, triplet_loss_name, distance_loss_name)
    if "huber" in distance_loss_name:
        best = _best_huber_distance(a_feature, pos_features, squared_d_dists, d_max_squared, f_max_squared)
    else:
        best = _best_distance(a_feature, pos_features, squared_d_dists, d_max_squared, f_max_squared)
    d_on = _sqdist(other_neg, neg_features) / f_max_squared
    second = torch.clamp(m2 + best[:, None] - d_on, min=0.0).max(dim=1).values.mean()
    return trip + second


def pairwise_distance_loss(anchor, positives, pairwise_squared_d_dists, d_max_squared, f_max_squared,
                           distance_loss_name="distance_loss"):
    """model/losses.py:627-646: all-pairs squared feature distances of [anchor, positives] (via :656-661) against the
    all-pairs squared metres, squared or Huber (labels=scaled_f, predictions=scaled_d), mean over everything."""
    anchor, positives, pairwise_squared_d_dists = map(_keep, (anchor, positives, pairwise_squared_d_dists))
    f = pairwise_squared_distances(torch.cat([anchor, positives], dim=1)) / f_max_squared
    d = pairwise_squared_d_dists / d_max_squared
    el = huber(f, d, reduce=False) if "huber" in distance_loss_name else (f - d) ** 2
    return el.mean(dim=2).mean(dim=1).mean(dim=0)


# --------------------------------------------------------------------------------------
# L6: logratio_loss (losses.py:125-135) -- literal broadcast, T=1 and P==N only
# --------------------------------------------------------------------------------------
def logratio_loss(a_feature, pos_features, neg_features, squared_pos_dists, squared_neg_dists):
    """model/losses.py:125-135 transcribed literally.

    tf.transpose without perm reverses all axes; with a [1,N] residual tensor that yields [N,1] and
    the division broadcasts to [N,P] (all pairs), whereas the [1,P,1] / [1,N,1]^T GPS ratio
    broadcasts to [1,P(=N),1] element-wise."""
    a_feature, pos_features, neg_features, squared_pos_dists, squared_neg_dists = map(
        _keep, (a_feature, pos_features, neg_features, squared_pos_dists, squared_neg_dists))
    pos_residuals = _sqdist(a_feature, pos_features)                          # [T,P]
    neg_residuals = _sqdist(a_feature, neg_features)                          # [T,N]
    rev = lambda x: x.permute(*reversed(range(x.dim())))
    feat_ratio = torch.log(pos_residuals / rev(neg_residuals))
    dist_ratio = torch.log(squared_pos_dists / rev(squared_neg_dists))
    squared_diffs = (feat_ratio - dist_ratio) ** 2
    return squared_diffs.mean(dim=1).mean(dim=1).mean(dim=0)


def logratio_loss_tuples(a_feature, pos_features, neg_features, squared_pos_dists, squared_neg_dists):
    """Tuple-mode batching: mean over tuples of the T=1 reference formula."""
    T = a_feature.shape[0]
    vals = [logratio_loss(a_feature[t:t + 1], pos_features[t:t + 1], neg_features[t:t + 1],
                          squared_pos_dists[t:t + 1], squared_neg_dists[t:t + 1]) for t in range(T)]
    return torch.stack(vals).mean()


# --------------------------------------------------------------------------------------
# D1: _pairwise_squared_distances (losses.py:656-661)
# --------------------------------------------------------------------------------------
def pairwise_squared_distances(features):
    """r - 2 X X^T + r^T, batched; no clamp, diagonal not forced to zero."""
    features = _keep(features)
    r = torch.einsum("aij,aij->ai", features, features)
    r = r.reshape(features.shape[0], -1, 1)
    batch_product = torch.einsum("aij,ajk->aik", features, features.permute(0, 2, 1))
    return r - 2 * batch_product + r.permute(0, 2, 1)


# --------------------------------------------------------------------------------------
# helpers for tests / bench
# --------------------------------------------------------------------------------------
def value_and_grad(fn, diff_args, *args, **kw):
    """Run ``fn(*diff_args, *args, **kw)`` in float64 and return (loss, [grads of diff_args])."""
    xs = [_t(a).clone().requires_grad_(True) for a in diff_args]
    loss = fn(*xs, *args, **kw)
    grads = torch.autograd.grad(loss, xs, allow_unused=True)
    return float(loss.detach()), [None if g is None else g.numpy() for g in grads]


def split_tuple(emb, P, N, other=False):
    """train/train.py:654: reshape(output,[T,S,-1]) split into [1,P,N(,1)] along axis 1."""
    q = emb[:, 0:1]
    pos = emb[:, 1:1 + P]
    neg = emb[:, 1 + P:1 + P + N]
    if other:
        return q, pos, neg, emb[:, 1 + P + N:2 + P + N]
    return q, pos, neg
