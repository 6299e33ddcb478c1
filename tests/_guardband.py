"""Guard-banded inputs for the mining comparisons of wms_loss / ms_loss.

The mining masks of model/losses.py:36-37 compare similarities against row thresholds; a pair that sits within float32
rounding of its threshold may legitimately flip between a float32 and a float64 evaluation and change the loss by
O(1e-2).  Parity tests therefore use batches in which NO pair lies within `band` of its threshold: the margins are
evaluated with the float64 oracle and the descriptors of offending pairs are re-drawn until the batch is clean.  With a
clean batch the kept-masks of the CUDA path must be IDENTICAL to the oracle's.  TEST INFRASTRUCTURE only."""
import numpy as np
import torch

from oracle import losses as ol


def mining_margins(mask_pos, mask_neg, emb64, eps=0.1):
    """|pos_mat - (max_val + eps)| where mask_pos > 0 and |neg_mat - (min_val - eps)| where mask_neg > 0 (losses.py:31-37),
    float64; entries the masks exclude are +inf."""
    e = ol.l2_normalize(torch.as_tensor(emb64, dtype=torch.float64), 1)
    s = torch.clamp(e @ e.T, min=0.0)
    pos, neg = s * mask_pos, s * mask_neg
    max_val = torch.amax(neg, dim=1, keepdim=True)
    tmp = torch.amax(pos, dim=1, keepdim=True)
    min_val = torch.amin((s - tmp) * mask_pos, dim=1, keepdim=True) + tmp
    inf = torch.full_like(s, float("inf"))
    mp = torch.where(mask_pos > 0, (pos - (max_val + eps)).abs(), inf)
    mn = torch.where(mask_neg > 0, (neg - (min_val - eps)).abs(), inf)
    return mp, mn


def wms_masks64(dist, d_alpha, d_beta, wfunction="exp"):
    d32 = torch.as_tensor(np.asarray(dist), dtype=torch.float32)
    mp, mn = ol.wms_masks(d32, d_alpha, d_beta, wfunction)
    mp = mp - torch.eye(d32.shape[0], dtype=torch.float32)
    return mp.double(), mn.double()


def ms_masks64(labels):
    lab = torch.as_tensor(np.asarray(labels)).reshape(-1, 1)
    adj = lab == lab.T
    return adj.double() - torch.eye(adj.shape[0], dtype=torch.float64), (~adj).double()


def guard_band(emb, masks, band=1e-4, seed=0, max_rounds=20, eps=0.1):
    """Re-draw (small fresh perturbation) the descriptor of the column of every pair within `band` of a mining
    threshold, until none is left.  `masks` is a list of (mask_pos, mask_neg) pairs that must ALL be clean (the same
    batch feeds wms and ms).  Returns (clean float32 descriptors, number of rounds, smallest remaining margin)."""
    rng = np.random.default_rng(seed)
    emb = np.array(emb, dtype=np.float32, copy=True)
    for rounds in range(max_rounds + 1):
        bad_cols, smallest = set(), float("inf")
        for mp_, mn_ in masks:
            mpos, mneg = mining_margins(mp_, mn_, emb.astype(np.float64), eps)
            m = torch.minimum(mpos, mneg)
            smallest = min(smallest, float(m.min()))
            ii, jj = torch.nonzero(m < band, as_tuple=True)
            bad_cols.update(int(j) for j in jj)
        if not bad_cols:
            return emb, rounds, smallest
        for j in sorted(bad_cols):
            emb[j] += (0.05 * np.abs(emb[j]).mean() * rng.standard_normal(emb.shape[1])).astype(np.float32)
    raise AssertionError(f"guard band not reached after {max_rounds} rounds (smallest margin {smallest:g})")
