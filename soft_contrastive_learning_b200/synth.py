"""Synthetic workloads of the shapes BASELINE.json names (no dataset / checkpoint is reachable offline).

Generators follow SURVEY.md section 8(d): tuples laid out as the reference's sampler emits them
(/root/reference/train/train.py:503,520: [anchor, P positives, N negatives(, other)]), GPS positions with
positives within MAX_POS_RADIUS=15 m (train.py:457) and mutually exclusive negatives >= MIN_NEG_RADIUS=15 m
(train.py:472-495), and the per-loss distance tensors built exactly as train.py:525-571 builds them.
NumPy only; used by tests, bench.py and smoke().
"""
from __future__ import annotations

import numpy as np

MAX_POS_RADIUS = 15.0
MIN_NEG_RADIUS = 15.0


def tuple_xy(rng, T, P, N, other=False, extent=1000.0):
    """xy [T,S,2] float64: anchor ~ U([0,extent]^2); positives within 15 m; negatives >= 15 m from anchor and each other."""
    S = 1 + P + N + (1 if other else 0)
    xy = np.empty((T, S, 2), dtype=np.float64)
    for t in range(T):
        a = rng.uniform(0, extent, size=2)
        xy[t, 0] = a
        r = rng.uniform(0, MAX_POS_RADIUS, size=P)
        th = rng.uniform(0, 2 * np.pi, size=P)
        xy[t, 1:1 + P, 0] = a[0] + r * np.cos(th)
        xy[t, 1:1 + P, 1] = a[1] + r * np.sin(th)
        placed = [a]
        n_need = N + (1 if other else 0)
        k = 0
        while k < n_need:
            c = rng.uniform(0, extent, size=2)
            if all(np.hypot(*(c - p)) >= MIN_NEG_RADIUS for p in placed):
                xy[t, 1 + P + k] = c
                placed.append(c)
                k += 1
    return xy


def tuple_descriptors(rng, T, P, N, D, other=False, pos_noise=0.7, dtype=np.float32):
    """emb [T,S,D]: z ~ N(0,I); positives = anchor + pos_noise*z so similarities straddle the mining thresholds."""
    S = 1 + P + N + (1 if other else 0)
    emb = rng.standard_normal((T, S, D))
    emb[:, 1:1 + P] = emb[:, 0:1] + pos_noise * emb[:, 1:1 + P]
    return emb.astype(dtype)


def pairwise_euclid(xy):
    """train.py:557-563 ('wms'): pairwise Euclidean metres over [anchor, pos..., neg...] -> [T,S,S]."""
    d = xy[:, :, None, :] - xy[:, None, :, :]
    return np.sqrt((d * d).sum(-1))


def anchor_sq_dists(xy, P):
    """train.py:529-534 ('anchor'): squared metres anchor -> each positive -> [T,P]."""
    d = xy[:, 1:1 + P] - xy[:, 0:1]
    return (d * d).sum(-1)


def logratio_sq_dists(xy, P, N):
    """train.py:564-571 ('logratio'): squared metres anchor->positives [T,P,1] and anchor->negatives [T,N,1]."""
    dp = xy[:, 1:1 + P] - xy[:, 0:1]
    dn = xy[:, 1 + P:1 + P + N] - xy[:, 0:1]
    return (dp * dp).sum(-1)[..., None], (dn * dn).sum(-1)[..., None]


def wms_batch(T=32, P=12, N=12, D=4096, seed=42, dtype=np.float32):
    """BASELINE config 1: T tuples of S=1+P+N descriptors with their [S,S] GPS distance matrices."""
    rng = np.random.default_rng(seed)
    xy = tuple_xy(rng, T, P, N)
    emb = tuple_descriptors(rng, T, P, N, D, dtype=dtype)
    dist = pairwise_euclid(xy).astype(dtype)
    return emb, dist, xy


def retrieval_problem(R, Q, D, seed=42, noise=0.5, dtype=np.float32, extent=10000.0):
    """BASELINE config 4/5 shape: db ~ N(0,1); queries = perturbed random db rows; xy ~ U([0,extent]^2)."""
    rng = np.random.default_rng(seed)
    db = rng.standard_normal((R, D), dtype=np.float32)
    src = rng.integers(0, R, size=Q)
    qry = db[src] + noise * rng.standard_normal((Q, D), dtype=np.float32)
    ref_xy = rng.uniform(0, extent, size=(R, 2))
    query_xy = ref_xy[src] + rng.normal(0, 3.0, size=(Q, 2))
    return db.astype(dtype), qry.astype(dtype), ref_xy, query_xy, src


def trajectory_layout(R, rng, stop_frac=0.05, stop_len=(100, 300)):
    """Frame index -> distance travelled along the route, in frames at cruising speed, for one traversal with stops.
    About ``stop_frac`` of the R frames belong to stops (vehicle standing: zero velocity for stop_len[0]..stop_len[1]
    consecutive frames).  Returns (s [R] float64, stopped [R] bool)."""
    stopped = np.zeros(R, dtype=bool)
    target = int(stop_frac * R)
    placed, guard = 0, 0
    while placed < target and guard < 10000:
        guard += 1
        n = int(rng.integers(stop_len[0], stop_len[1] + 1))
        t0 = int(rng.integers(0, max(1, R - n)))
        if stopped[max(0, t0 - 1):t0 + n + 1].any():
            continue
        stopped[t0:t0 + n] = True
        placed += n
    s = np.cumsum(~stopped).astype(np.float64)
    return s, stopped


def trajectory_rows(s, stopped, anchors, seg_len, frame_noise, stop_noise, noise):
    """Descriptors of frames at route positions ``s``: great-circle interpolation between consecutive random anchor
    descriptors (cos/sin weights keep every dimension at unit variance: PCA-whitened statistics), plus per-frame noise
    (``stop_noise`` while standing, ``frame_noise`` otherwise).  Works on NumPy arrays and on torch tensors alike."""
    seg = (s // seg_len)
    theta = (s - seg * seg_len) * (np.pi / 2 / seg_len)
    if type(s).__module__.startswith("torch"):
        import torch
        seg = seg.long()
        sig = torch.where(stopped, torch.full_like(theta, stop_noise), torch.full_like(theta, frame_noise))
        x = torch.cos(theta)[:, None].float() * anchors[seg] + torch.sin(theta)[:, None].float() * anchors[seg + 1]
        return x + sig[:, None].float() * noise
    seg = seg.astype(np.int64)
    sig = np.where(stopped, stop_noise, frame_noise)
    x = np.cos(theta)[:, None] * anchors[seg] + np.sin(theta)[:, None] * anchors[seg + 1]
    return (x + sig[:, None] * noise).astype(np.float32)


def trajectory_problem(R, Q, D, seed=42, seg_len=64, frame_noise=0.05, stop_noise=2e-3, stop_frac=0.05,
                       stop_len=(100, 300), query_noise=0.05):
    """Clustered retrieval data as a real traversal produces it (SURVEY 8d `db = f(xy) + noise`):
    the database holds the descriptors of R CONSECUTIVE frames of one drive -- neighbouring frames are near-duplicates
    (squared distance ~ D (pi/2/seg_len)^2 j^2 for frames j apart) and at the stops hundreds of frames are identical up
    to sensor noise, far inside the fp16 rounding bound of the tensor pass.  Queries are frames of a second drive:
    perturbed database frames, about ``stop_frac`` of them on a stop.
    Returns (db [R,D] f32, queries [Q,D] f32, info dict with 'src', 'stopped', 's')."""
    rng = np.random.default_rng(seed)
    s, stopped = trajectory_layout(R, rng, stop_frac, stop_len)
    n_anchor = int(s[-1] // seg_len) + 2
    anchors = rng.standard_normal((n_anchor, D)).astype(np.float32)
    db = trajectory_rows(s, stopped, anchors, seg_len, frame_noise, stop_noise, rng.standard_normal((R, D)).astype(np.float32))
    src = rng.integers(0, R, size=Q)
    qry = (db[src] + query_noise * rng.standard_normal((Q, D))).astype(np.float32)
    return db, qry, {"src": src, "stopped": stopped, "s": s}


def netvlad_problem(B=2, H=3, W=4, C=512, K=64, Dout=128, seed=42):
    """BASELINE config 2 shape (scaled by the caller): conv5 maps, assignment weights, centres, PCA (V, m, var)."""
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, H, W, C)).astype(np.float32)
    aw = (0.05 * rng.standard_normal((C, K))).astype(np.float32)
    cc = (0.05 * rng.standard_normal((C, K))).astype(np.float32)
    Din = C * K
    # random orthonormal rows via QR of a thin Gaussian (Dout << Din)
    g = rng.standard_normal((Din, Dout))
    q, _ = np.linalg.qr(g)
    V = q.T.astype(np.float32)
    m = (0.01 * rng.standard_normal(Din)).astype(np.float32)
    var = rng.uniform(0.5, 2.0, size=Dout).astype(np.float32)
    return x, aw, cc, V, m, var


def pca_features(n, D, rank=40, seed=42, dtype=np.float64):
    """Features for the PCA fit: `rank` orthogonal directions with geometrically decaying strength over
    small isotropic noise and a non-zero mean, so the leading components are well separated."""
    rng = np.random.default_rng(seed)
    basis = np.linalg.qr(rng.standard_normal((D, rank)))[0].T
    strength = 10.0 * 0.8 ** np.arange(rank)
    x = (rng.standard_normal((n, rank)) * strength) @ basis + 0.05 * rng.standard_normal((n, D)) + rng.standard_normal(D)
    return x.astype(dtype)
