"""Oracle (TEST INFRASTRUCTURE): NetVLAD aggregation head and PCA-whitening projection.

N1 -- ``layers.netVLAD(x, 64)`` is called at /root/reference/model/nets.py:66-67 (and
model/grad_nets.py:66-67) but its body lives in the un-vendored, un-pinned third-party repo
uzh-rpg/netvlad_tf_open (README.md:10): PARITY UNPINNED.  Restated from that repo's published
``netvlad_tf/layers.py``:

    s = conv2d(inputs, K, 1x1, use_bias=False)           # 'assignment/kernel' [1,1,D,K]
    a = softmax(s)                                       # over clusters
    v = sum_{h,w} a[..., None, :] * (inputs[..., None] + C)   # 'cluster_centers' [1,1,1,D,K], stored negated
    v = transpose(v, [0,2,1]); v /= sqrt(sum(v^2, -1) + 1e-12)  # intra-normalisation per cluster
    v = transpose(v, [0,2,1]); v = flatten(v)            # index d*K + k
    v /= sqrt(sum(v^2, -1) + 1e-12)

preceded, in-repo, by ``x = tf.nn.l2_normalize(x, axis=-1)`` (nets.py:66).

P1 -- the PCA op of train/train.py:646-652:  y = matmul(x - m, v, adjoint_b=True) / sqrt(var);
its evaluation twin is sklearn ``PCA(whiten=True).transform`` (evaluation/top-n.py:74-77).
"""
from __future__ import annotations

import numpy as np
import torch

from .losses import _keep, l2_normalize


def netvlad_head(conv5, assign_w, centers, pre_l2norm=True):
    """conv5 [B,h,w,D] (NHWC) or [B,HW,D]; assign_w [D,K]; centers [D,K] -> [B, D*K]."""
    conv5, assign_w, centers = map(_keep, (conv5, assign_w, centers))
    B = conv5.shape[0]
    D = conv5.shape[-1]
    x = conv5.reshape(B, -1, D)
    if pre_l2norm:
        x = l2_normalize(x, -1)                                   # nets.py:66
    s = x @ assign_w                                              # 1x1 conv, no bias
    a = torch.softmax(s, dim=-1)                                  # [B,HW,K]
    # sum_n a[n,k] * (x[n,d] + C[d,k])
    v = torch.einsum("bnk,bnd->bdk", a, x) + centers[None] * a.sum(dim=1)[:, None, :]
    v = v / torch.sqrt((v * v).sum(dim=1, keepdim=True) + 1e-12)  # intra-norm over d, per cluster
    v = v.reshape(B, -1)                                          # [B, D*K], index d*K+k
    v = v / torch.sqrt((v * v).sum(dim=1, keepdim=True) + 1e-12)
    return v


def pca_project(x, v, m, var):
    """train/train.py:650-651: (x - m) @ v^T / sqrt(var).  x [B,Din], v [Dout,Din], m [Din], var [Dout]."""
    x, v, m, var = map(_keep, (x, v, m, var))
    return ((x - m) @ v.T) / torch.sqrt(var)


def pca_fit(x, n_components):
    """``PCA(whiten=True, n_components=d).fit(x)`` (evaluation/top-n.py:74-75) restated in float64 NumPy after
    scikit-learn's exact solver (third-party, version unpinned by the reference; ``PCA._fit_full``): centre by the
    column mean, thin SVD, ``svd_flip(u_based_decision=False)`` (largest-magnitude entry of every row of V^T made
    positive), ``explained_variance_ = S**2 / (n - 1)``.  Returns (components_ [d,D], mean_ [D], explained_variance_ [d])."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    mean = x.mean(axis=0)
    _, S, Vt = np.linalg.svd(x - mean, full_matrices=False)
    idx = np.argmax(np.abs(Vt), axis=1)
    Vt = Vt * np.sign(Vt[np.arange(Vt.shape[0]), idx])[:, None]
    d = int(n_components)
    return Vt[:d], mean, (S[:d] ** 2) / (n - 1)


def sklearn_pca_params(pca):
    """Map a fitted sklearn PCA(whiten=True) (top-n.py:74-75) onto the (v, m, var) of the train-time op."""
    return (np.asarray(pca.components_), np.asarray(pca.mean_), np.asarray(pca.explained_variance_))
