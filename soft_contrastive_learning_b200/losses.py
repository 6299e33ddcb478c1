"""Metric-learning losses of the reference, by the reference's own names, on B200.

Host-side mirror of /root/reference/model/losses.py and the loss-by-name dispatch of
/root/reference/train/train.py:700-855.  Every function keeps the reference's signature and argument meaning;
inputs are torch CUDA tensors (or NumPy arrays, which are copied to the current device), the result is a
0-d tensor that carries the analytic gradient w.r.t. the descriptors (what ``optimizer.minimize`` obtained
from TF autodiff, train.py:874-878).  All arithmetic happens in libscl_b200.so (csrc/*.cu); torch only owns
memory and streams.  There is no CPU or eager fallback.

Fused fast path: ``*_value_and_grad`` return ``(loss, d loss / d embeddings)`` from ONE kernel launch
(forward and backward share the shared-memory-resident descriptors) without building an autograd graph.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import MsParams, TupleParams, check, lib


# ----------------------------------------------------------------------------------------------
# plumbing
# ----------------------------------------------------------------------------------------------
def _dev():
    if not torch.cuda.is_available():
        raise _lib.SclError("soft_contrastive_learning_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _f32(x):
    """Contiguous float32 tensor on the current CUDA device (NumPy / host tensors are copied over)."""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if not isinstance(x, torch.Tensor):
        x = torch.as_tensor(x)
    if not x.is_cuda:
        x = x.to(_dev(), non_blocking=True)
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous()


def _p(t):
    return C.c_void_p(0 if t is None else t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ws(nbytes, device):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def _ms_params(d_alpha=0.0, d_beta=1.0, alpha=2.0, beta=50.0, lamb=1.0, eps=0.1, ms_mining=True, wfunction="exp",
               sumfunction="ms"):
    # the reference falls through to 'exp' for any unknown wfunction (losses.py:17) and computes nothing for an
    # unknown sumfunction; unknown names are rejected here
    if wfunction not in _lib.WF:
        wfunction = "exp"
    if sumfunction not in _lib.SUMF:
        raise ValueError(f"sumfunction must be 'ms' or 'plain', got {sumfunction!r}")
    return MsParams(float(d_alpha), float(d_beta), float(alpha), float(beta), float(lamb), float(eps),
                    int(bool(ms_mining)), _lib.WF[wfunction], _lib.SUMF[sumfunction])


# ----------------------------------------------------------------------------------------------
# W1 / W2 raw calls
# ----------------------------------------------------------------------------------------------
def _wms_tuple_raw(emb3, dist3, params, need_grad=True, want_kept=False, want_per_tuple=False):
    T, S, D = emb3.shape
    dev = emb3.device
    L = lib()
    nbytes = C.c_size_t()
    check(L.scl_wms_tuple_workspace_bytes(T, S, D, C.byref(nbytes)), "scl_wms_tuple_workspace_bytes")
    ws = _ws(nbytes.value, dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    grad = torch.empty_like(emb3) if need_grad else None
    kept = torch.empty((T, S, 2), dtype=torch.int32, device=dev) if want_kept else None
    per = torch.empty(T, dtype=torch.float32, device=dev) if want_per_tuple else None
    check(L.scl_wms_tuple_fwd_bwd(_p(emb3), _p(dist3), T, S, D, C.byref(params), _p(loss), _p(per), _p(grad), _p(kept),
                                  _p(ws), ws.numel(), _stream()), "scl_wms_tuple_fwd_bwd")
    return loss, grad, kept, per


def _flat_raw(emb2, dist2, labels, params, need_grad=True, want_kept=False):
    B, D = emb2.shape
    dev = emb2.device
    L = lib()
    nbytes = C.c_size_t()
    check(L.scl_ms_flat_workspace_bytes(B, D, C.byref(nbytes)), "scl_ms_flat_workspace_bytes")
    ws = _ws(nbytes.value, dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    grad = torch.empty_like(emb2) if need_grad else None
    kept = torch.empty((2, B, B), dtype=torch.uint8, device=dev) if want_kept else None
    if dist2 is not None:
        check(L.scl_wms_flat_fwd_bwd(_p(emb2), _p(dist2), B, D, C.byref(params), _p(loss), _p(grad), _p(kept), _p(ws),
                                     ws.numel(), _stream()), "scl_wms_flat_fwd_bwd")
    else:
        check(L.scl_ms_flat_fwd_bwd(_p(emb2), _p(labels), B, D, C.byref(params), _p(loss), _p(grad), _p(kept), _p(ws),
                                    ws.numel(), _stream()), "scl_ms_flat_fwd_bwd")
    return loss, grad, kept


_TUPLE_MAX_S = 32


def _wms_dispatch(distances, embeddings, params, need_grad, want_kept=False):
    """Route to tuple mode ([T,S,S] distances, S<=32) or flat mode ([B,B] distances)."""
    dist = _f32(distances)
    emb = _f32(embeddings)
    if dist.dim() == 3:
        T, S, _ = dist.shape
        emb3 = emb.reshape(T, S, -1)
        if S > _TUPLE_MAX_S:
            raise ValueError("tuple-mode wms_loss supports at most 32 descriptors per tuple")
        loss, grad, kept, _ = _wms_tuple_raw(emb3, dist, params, need_grad, want_kept)
        if grad is not None:
            grad = grad.reshape(emb.shape)
        return loss, grad, kept
    if dist.dim() != 2 or dist.shape[0] != dist.shape[1]:
        raise ValueError("distances must be [B,B] or [T,S,S]")
    B = dist.shape[0]
    emb2 = emb.reshape(B, -1)
    if B <= _TUPLE_MAX_S and emb2.shape[1] % 4 == 0:
        loss, grad, kept, _ = _wms_tuple_raw(emb2[None], dist[None], params, need_grad, want_kept)
        if grad is not None:
            grad = grad.reshape(emb.shape)
        return loss, grad, kept
    loss, grad, kept = _flat_raw(emb2, dist, None, params, need_grad, want_kept)
    if grad is not None:
        grad = grad.reshape(emb.shape)
    return loss, grad, kept


class _PrecomputedGrad(torch.autograd.Function):
    """loss with its gradient already computed by the fused kernel: backward is one scale."""

    @staticmethod
    def forward(ctx, emb, loss, grad):
        ctx.save_for_backward(grad)
        return loss.reshape(()).clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return grad * g, None, None


def _attach(emb_in, loss, grad):
    if isinstance(emb_in, torch.Tensor) and emb_in.requires_grad:
        if grad.shape != emb_in.shape:
            grad = grad.reshape(emb_in.shape)
        return _PrecomputedGrad.apply(emb_in, loss, grad.to(emb_in.dtype))
    return loss.reshape(())


def _to_host(loss, grad, like):
    if isinstance(like, np.ndarray):
        return float(loss.item()), (None if grad is None else grad.cpu().numpy())
    return loss.reshape(()), grad


# ----------------------------------------------------------------------------------------------
# public: the reference's names
# ----------------------------------------------------------------------------------------------
def wms_loss(distances, embeddings, d_alpha, d_beta, alpha=2.0, beta=50.0, lamb=1.0, eps=0.1, ms_mining=True,
             wfunction="exp", sumfunction="ms"):
    """Weighted multi-similarity loss -- model/losses.py:5-60 (call train/train.py:852).

    ``distances``: pairwise Euclidean GPS distances in metres, [B,B], or [T,S,S] as train.py:684-686 feeds it
    (then ``embeddings`` is [T*S,D] or [T,S,D] and the result is the mean over tuples of the per-tuple loss;
    for T=1 this is exactly the reference's value)."""
    params = _ms_params(d_alpha, d_beta, alpha, beta, lamb, eps, ms_mining, wfunction, sumfunction)
    need = isinstance(embeddings, torch.Tensor) and embeddings.requires_grad
    loss, grad, _ = _wms_dispatch(distances, embeddings, params, need_grad=need)
    return _attach(embeddings, loss, grad) if need else loss.reshape(())


def wms_loss_value_and_grad(distances, embeddings, d_alpha, d_beta, alpha=2.0, beta=50.0, lamb=1.0, eps=0.1,
                            ms_mining=True, wfunction="exp", sumfunction="ms", return_kept=False):
    """Fused forward+backward: (loss, d loss/d embeddings[, kept-pair masks]).  NumPy in -> NumPy out."""
    params = _ms_params(d_alpha, d_beta, alpha, beta, lamb, eps, ms_mining, wfunction, sumfunction)
    loss, grad, kept = _wms_dispatch(distances, embeddings, params, need_grad=True, want_kept=return_kept)
    out = _to_host(loss, grad, embeddings)
    return out + (kept,) if return_kept else out


def _labels_i32(labels, dev):
    lab = np.asarray(labels.detach().cpu() if isinstance(labels, torch.Tensor) else labels).reshape(-1)
    _, inv = np.unique(lab, return_inverse=True)          # equality classes only (losses.py:89)
    return torch.from_numpy(inv.astype(np.int32)).to(dev)


def ms_loss(labels, embeddings, alpha=2.0, beta=50.0, lamb=1.0, eps=0.1, ms_mining=True):
    """Multi-similarity loss -- model/losses.py:76-122 (call train/train.py:821-827)."""
    need = isinstance(embeddings, torch.Tensor) and embeddings.requires_grad
    emb = _f32(embeddings)
    emb2 = emb.reshape(-1, emb.shape[-1])
    params = _ms_params(0.0, 1.0, alpha, beta, lamb, eps, ms_mining)
    loss, grad, _ = _flat_raw(emb2, None, _labels_i32(labels, emb.device), params, need_grad=need)
    return _attach(embeddings, loss, grad) if need else loss.reshape(())


def ms_loss_value_and_grad(labels, embeddings, alpha=2.0, beta=50.0, lamb=1.0, eps=0.1, ms_mining=True,
                           return_kept=False):
    emb = _f32(embeddings)
    emb2 = emb.reshape(-1, emb.shape[-1])
    params = _ms_params(0.0, 1.0, alpha, beta, lamb, eps, ms_mining)
    loss, grad, kept = _flat_raw(emb2, None, _labels_i32(labels, emb.device), params, True, return_kept)
    out = _to_host(loss, grad.reshape(emb.shape), embeddings)
    return out + (kept,) if return_kept else out


def ms_labels(tuples_per_batch, positives_per_tuple, negatives_per_tuple):
    """Class labels exactly as train/train.py:822-826 builds them."""
    one = np.concatenate((np.zeros(1 + positives_per_tuple), np.arange(negatives_per_tuple) + 1))
    all_labels = one
    for batch in range(1, tuples_per_batch):
        all_labels = np.concatenate((all_labels, one + batch * (negatives_per_tuple + 1)))
    return all_labels


# ---------------- triplet family ----------------
def _tuple_raw(emb3, P, N, sq_d, tp, need_grad=True):
    T, S, D = emb3.shape
    dev = emb3.device
    L = lib()
    nbytes = C.c_size_t()
    check(L.scl_tuple_loss_workspace_bytes(T, P, N, D, C.byref(nbytes)), "scl_tuple_loss_workspace_bytes")
    ws = _ws(nbytes.value, dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    grad = torch.empty_like(emb3) if need_grad else None
    check(L.scl_tuple_loss_fwd_bwd(_p(emb3), T, P, N, D, _p(sq_d), C.byref(tp), _p(loss), _p(grad), _p(ws), ws.numel(),
                                   _stream()), "scl_tuple_loss_fwd_bwd")
    return loss, grad


def _cat_tuple(parts):
    parts = [_f32(p) for p in parts]
    parts = [p if p.dim() == 3 else p.reshape(p.shape[0], -1, p.shape[-1]) for p in parts]
    return torch.cat(parts, dim=1).contiguous()


class _TupleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tp, sq_d, sizes, *parts):
        emb3 = _cat_tuple([p.detach() for p in parts])
        P, N = sizes
        loss, grad = _tuple_raw(emb3, P, N, sq_d, tp, need_grad=True)
        ctx.splits = [p.shape[1] if p.dim() == 3 else 1 for p in parts]
        ctx.shapes = [p.shape for p in parts]
        ctx.save_for_backward(grad)
        return loss.reshape(()).clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        outs = torch.split(grad * g, ctx.splits, dim=1)
        return (None, None, None) + tuple(o.reshape(s) for o, s in zip(outs, ctx.shapes))


def _tuple_loss(kind, parts, m1, m2=0.0, lam=0.0, sq_d=None, d_max=1.0, f_max=1.0, dist_term="none"):
    tp = TupleParams(_lib.TUPLE_KIND[kind], _lib.DIST_TERM[dist_term], float(m1), float(m2), float(lam), float(d_max),
                     float(f_max))
    tparts = [p if isinstance(p, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(p)) for p in parts]
    tparts = [p if p.is_cuda else p.to(_dev()) for p in tparts]
    P = tparts[1].shape[1]
    N = tparts[2].shape[1]
    sq = None if sq_d is None else _f32(sq_d).reshape(-1, P)
    if any(p.requires_grad for p in tparts):
        return _TupleFn.apply(tp, sq, (P, N), *tparts)
    loss, _ = _tuple_raw(_cat_tuple(tparts), P, N, sq, tp, need_grad=False)
    return loss.reshape(())


def triplet_loss(q_vec, pos_vecs, neg_vecs, margin):
    """pointnetvlad_cls.triplet_loss (train/train.py:701): mean_t sum_n max(m + min_p|q-p|^2 - |q-n|^2, 0)."""
    return _tuple_loss("triplet_loss", (q_vec, pos_vecs, neg_vecs), margin)


def lazy_triplet_loss(q_vec, pos_vecs, neg_vecs, margin):
    """pointnetvlad_cls.lazy_triplet_loss (train/train.py:703)."""
    return _tuple_loss("lazy_triplet_loss", (q_vec, pos_vecs, neg_vecs), margin)


def quadruplet_loss(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2):
    """pointnetvlad_cls.quadruplet_loss (train/train.py:707-708)."""
    return _tuple_loss("quadruplet_loss", (q_vec, pos_vecs, neg_vecs, other_neg), m1, m2)


def lazy_quadruplet_loss(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2):
    """pointnetvlad_cls.lazy_quadruplet_loss (train/train.py:710-712)."""
    return _tuple_loss("lazy_quadruplet_loss", (q_vec, pos_vecs, neg_vecs, other_neg), m1, m2)


def evil_triplet_loss(q_vec, pos_vecs, neg_vecs, margin):
    """model/losses.py:63-73."""
    return _tuple_loss("evil_triplet_loss", (q_vec, pos_vecs, neg_vecs), margin)


def evil_quadruplet_loss(q_vec, pos_vecs, neg_vecs, other_neg, m1, m2):
    """model/losses.py:197-214."""
    return _tuple_loss("evil_quadruplet_loss", (q_vec, pos_vecs, neg_vecs, other_neg), m1, m2)


def distance_triplet_loss(a_feature, pos_features, neg_features, margin, lam, squared_d_dists, d_max_squared,
                          f_max_squared, triplet_loss_name="triplet_loss", distance_loss_name="huber_distance_loss"):
    """model/losses.py:239-264 (calls train/train.py:719-747): triplet + lam * (Huber) distance term."""
    if triplet_loss_name not in ("triplet_loss", "lazy_triplet_loss"):
        raise AttributeError(f"module 'pointnetvlad_cls' has no attribute {triplet_loss_name!r}")
    term = "huber_distance_loss" if "huber" in distance_loss_name else "distance_loss"     # losses.py:255
    return _tuple_loss(triplet_loss_name, (a_feature, pos_features, neg_features), margin, 0.0, lam, squared_d_dists,
                       d_max_squared, f_max_squared, term)


def distance_quadruplet_loss(a_feature, pos_features, neg_features, other_neg, m1, m2, lam, squared_d_dists,
                             d_max_squared, f_max_squared, triplet_loss_name="triplet_loss",
                             distance_loss_name="huber_distance_loss"):
    """model/losses.py:267-307 (calls train/train.py:729-763): distance_triplet_loss + the distance-term second hinge."""
    if triplet_loss_name not in ("triplet_loss", "lazy_triplet_loss"):
        raise AttributeError(f"module 'pointnetvlad_cls' has no attribute {triplet_loss_name!r}")
    kind = "distance_lazy_quadruplet_loss" if triplet_loss_name == "lazy_triplet_loss" else "distance_quadruplet_loss"
    term = "huber_distance_loss" if "huber" in distance_loss_name else "distance_loss"     # losses.py:285,292
    return _tuple_loss(kind, (a_feature, pos_features, neg_features, other_neg), m1, m2, lam, squared_d_dists,
                       d_max_squared, f_max_squared, term)


def _pairwise_distance_raw(emb3, sq_d3, d_max_squared, f_max_squared, huber, need_grad=True):
    T, n, D = emb3.shape
    L = lib()
    nbytes = C.c_size_t()
    check(L.scl_wms_tuple_workspace_bytes(T, n, D, C.byref(nbytes)), "scl_wms_tuple_workspace_bytes")
    ws = _ws(nbytes.value, emb3.device)
    loss = torch.empty(1, dtype=torch.float32, device=emb3.device)
    grad = torch.empty_like(emb3) if need_grad else None
    check(L.scl_pairwise_distance_loss_fwd_bwd(_p(emb3), _p(sq_d3), T, n, D, float(d_max_squared), float(f_max_squared),
                                               int(huber), _p(loss), _p(grad), _p(ws), ws.numel(), _stream()),
          "scl_pairwise_distance_loss_fwd_bwd")
    return loss, grad


class _PairwiseDistanceFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sq_d, consts, anchor, positives):
        emb3 = _f32(torch.cat([anchor.detach(), positives.detach()], dim=1))
        loss, grad = _pairwise_distance_raw(emb3, sq_d, *consts, need_grad=True)
        ctx.save_for_backward(grad)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        grad = grad * g
        return None, None, grad[:, :1], grad[:, 1:]


def pairwise_distance_loss(anchor, positives, pairwise_squared_d_dists, d_max_squared, f_max_squared,
                           distance_loss_name="distance_loss"):
    """model/losses.py:627-646: anchor [T,1,D], positives [T,P,D], pairwise_squared_d_dists [T,1+P,1+P] (squared
    metres, DISTANCE_TYPE 'pairwise', train.py:535-537) -> scalar; squared or Huber element (name contains 'huber')."""
    parts = [p if isinstance(p, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(p)) for p in (anchor, positives)]
    parts = [p if p.is_cuda else p.to(_dev()) for p in parts]
    n = 1 + parts[1].shape[1]
    sq = _f32(pairwise_squared_d_dists).reshape(-1, n, n)
    consts = (float(d_max_squared), float(f_max_squared), 1 if "huber" in distance_loss_name else 0)
    if any(p.requires_grad for p in parts):
        return _PairwiseDistanceFn.apply(sq, consts, *parts)
    loss, _ = _pairwise_distance_raw(_f32(torch.cat(parts, dim=1)), sq, *consts, need_grad=False)
    return loss.reshape(())


def pairwise_distance_loss_value_and_grad(features, pairwise_squared_d_dists, d_max_squared=225.0, f_max_squared=2.0,
                                          distance_loss_name="distance_loss"):
    """Fused path on [T, 1+P, D] features (anchor first)."""
    emb3 = _f32(features)
    sq = _f32(pairwise_squared_d_dists).reshape(emb3.shape[0], emb3.shape[1], emb3.shape[1])
    loss, grad = _pairwise_distance_raw(emb3, sq, d_max_squared, f_max_squared, 1 if "huber" in distance_loss_name else 0)
    return _to_host(loss, grad, features)


def tuple_loss_value_and_grad(name, output, tuples_per_batch, positives_per_tuple, negatives_per_tuple, m1=0.1, m2=0.2,
                              lam=0.5, squared_d_dists=None, d_max_squared=225.0, f_max_squared=2.0,
                              distance_loss_name="none"):
    """Fused path on the un-split network output [T*S, D] (train.py:654 layout).  ``name`` is a TUPLE_KIND key."""
    emb = _f32(output)
    other = 1 if "quadruplet" in name else 0
    S = 1 + positives_per_tuple + negatives_per_tuple + other
    emb3 = emb.reshape(tuples_per_batch, S, -1)
    tp = TupleParams(_lib.TUPLE_KIND[name], _lib.DIST_TERM[distance_loss_name], float(m1), float(m2), float(lam),
                     float(d_max_squared), float(f_max_squared))
    sq = None if squared_d_dists is None else _f32(squared_d_dists).reshape(tuples_per_batch, positives_per_tuple)
    loss, grad = _tuple_raw(emb3, positives_per_tuple, negatives_per_tuple, sq, tp, need_grad=True)
    return _to_host(loss, grad.reshape(emb.shape), output)


# ---------------- logratio ----------------
def _logratio_raw(emb3, P, N, sq_pos, sq_neg, strict, need_grad=True):
    T, S, D = emb3.shape
    dev = emb3.device
    L = lib()
    ws = _ws(256 + 4 * (T + 8), dev)
    loss = torch.empty(1, dtype=torch.float32, device=dev)
    grad = torch.empty_like(emb3) if need_grad else None
    check(L.scl_logratio_fwd_bwd(_p(emb3), T, P, N, D, _p(sq_pos), _p(sq_neg), int(strict), _p(loss), _p(grad), _p(ws),
                                 ws.numel(), _stream()), "scl_logratio_fwd_bwd")
    return loss, grad


class _LogratioFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, sq_pos, sq_neg, strict, a, pos, neg):
        emb3 = _cat_tuple([a.detach(), pos.detach(), neg.detach()])
        loss, grad = _logratio_raw(emb3, pos.shape[1], neg.shape[1], sq_pos, sq_neg, strict)
        ctx.splits = [1, pos.shape[1], neg.shape[1]]
        ctx.save_for_backward(grad)
        return loss.reshape(()).clone()

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return (None, None, None) + tuple(torch.split(grad * g, ctx.splits, dim=1))


def logratio_loss(a_feature, pos_features, neg_features, squared_pos_dists, squared_neg_dists, strict_reference=True):
    """model/losses.py:125-135 (call train/train.py:854-855).  Inputs [T,1,D], [T,P,D], [T,N,D], [T,P,1], [T,N,1];
    T>1 averages the reference's T=1 formula over tuples.  ``strict_reference`` keeps the reference's broadcast
    (needs P == N); False uses the all-pairs GPS ratio."""
    parts = [p if isinstance(p, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(p))
             for p in (a_feature, pos_features, neg_features)]
    parts = [p if p.is_cuda else p.to(_dev()) for p in parts]
    P, N = parts[1].shape[1], parts[2].shape[1]
    sp = _f32(squared_pos_dists).reshape(-1, P)
    sn = _f32(squared_neg_dists).reshape(-1, N)
    if any(p.requires_grad for p in parts):
        return _LogratioFn.apply(sp, sn, strict_reference, *parts)
    loss, _ = _logratio_raw(_cat_tuple(parts), P, N, sp, sn, strict_reference, need_grad=False)
    return loss.reshape(())


def logratio_loss_value_and_grad(output, tuples_per_batch, positives_per_tuple, negatives_per_tuple, squared_pos_dists,
                                 squared_neg_dists, strict_reference=True):
    emb = _f32(output)
    emb3 = emb.reshape(tuples_per_batch, 1 + positives_per_tuple + negatives_per_tuple, -1)
    sp = _f32(squared_pos_dists).reshape(tuples_per_batch, positives_per_tuple)
    sn = _f32(squared_neg_dists).reshape(tuples_per_batch, negatives_per_tuple)
    loss, grad = _logratio_raw(emb3, positives_per_tuple, negatives_per_tuple, sp, sn, strict_reference)
    return _to_host(loss, grad.reshape(emb.shape), output)


# ---------------- D1 ----------------
def _pairwise_squared_distances(features):
    """model/losses.py:656-661: [T,n,D] -> [T,n,n], r_i - 2 x_i.x_j + r_j (forward only, as used by :627-646)."""
    x = _f32(features)
    T, n, D = x.shape
    out = torch.empty((T, n, n), dtype=torch.float32, device=x.device)
    check(lib().scl_pairwise_sqdist(_p(x), T, n, D, _p(out), _stream()), "scl_pairwise_sqdist")
    return out.cpu().numpy() if isinstance(features, np.ndarray) else out


pairwise_squared_distances = _pairwise_squared_distances


# ----------------------------------------------------------------------------------------------
# loss-by-name, the boundary train.py:700-855 defines
# ----------------------------------------------------------------------------------------------
LOSS_NAMES = ("triplet", "lazy_triplet", "evil_triplet", "quadruplet", "lazy_quadruplet", "evil_quadruplet",
              "distance_triplet", "distance_lazy_triplet", "huber_distance_triplet", "huber_distance_lazy_triplet",
              "distance_quadruplet", "distance_lazy_quadruplet", "huber_distance_quadruplet",
              "huber_distance_lazy_quadruplet", "ms_loss", "wms", "logratio")


def split_outputs(output, tuples_per_batch, positives_per_tuple, negatives_per_tuple, other=False):
    """train.py:654: tf.split(tf.reshape(output, [T, sum(tuple_shape), -1]), tuple_shape, 1)."""
    shape = [1, positives_per_tuple, negatives_per_tuple] + ([1] if other else [])
    o3 = output.reshape(tuples_per_batch, sum(shape), -1)
    return torch.split(o3, shape, dim=1)


def get_loss(name):
    """Return ``f(output, distances, cfg) -> loss`` for a ``--loss`` name of train/train.py:1222-1469.

    ``cfg`` carries the globals the reference reads: TUPLES_PER_BATCH, POSITIVES_PER_TUPLE, NEGATIVES_PER_TUPLE
    (the latter already decremented for quadruplet losses, train.py:589-592), MARGIN_1, MARGIN_2, LAM, ALPHA, BETA,
    MAX_POS_RADIUS, MSMINING, WFUNCTION, SUMFUNCTION."""
    if name not in LOSS_NAMES:
        raise KeyError(f"loss {name!r} is outside the hot path built here; available: {LOSS_NAMES}")

    def f(output, distances=None, cfg=None):
        c = dict(TUPLES_PER_BATCH=1, POSITIVES_PER_TUPLE=12, NEGATIVES_PER_TUPLE=12, MARGIN_1=0.1, MARGIN_2=0.2, LAM=0.5,
                 ALPHA=0.8, BETA=15.0, MAX_POS_RADIUS=15.0, MSMINING=True, WFUNCTION="exp", SUMFUNCTION="ms")
        c.update(cfg or {})
        T, P, N = c["TUPLES_PER_BATCH"], c["POSITIVES_PER_TUPLE"], c["NEGATIVES_PER_TUPLE"]
        d_max_squared = float(c["MAX_POS_RADIUS"]) ** 2                       # train.py:695
        f_max_squared = 2.0                                                   # train.py:696
        quad = "quadruplet" in name
        outs = split_outputs(output, T, P, N, other=quad)
        if name == "triplet":                                                 # train.py:700-701
            return triplet_loss(outs[0], outs[1], outs[2], c["MARGIN_1"])
        if name == "lazy_triplet":
            return lazy_triplet_loss(outs[0], outs[1], outs[2], c["MARGIN_1"])
        if name == "evil_triplet":
            return evil_triplet_loss(outs[0], outs[1], outs[2], c["MARGIN_1"])
        if name == "quadruplet":
            return quadruplet_loss(outs[0], outs[1], outs[2], outs[3], c["MARGIN_1"], c["MARGIN_2"])
        if name == "lazy_quadruplet":
            return lazy_quadruplet_loss(outs[0], outs[1], outs[2], outs[3], c["MARGIN_1"], c["MARGIN_2"])
        if name == "evil_quadruplet":
            return evil_quadruplet_loss(outs[0], outs[1], outs[2], outs[3], c["MARGIN_1"], c["MARGIN_2"])
        if name in ("distance_triplet", "distance_lazy_triplet", "huber_distance_triplet",
                    "huber_distance_lazy_triplet"):                           # train.py:719-747
            trip = "lazy_triplet_loss" if "lazy" in name else "triplet_loss"
            dl = "huber_distance_loss" if "huber" in name else "distance_loss"
            return distance_triplet_loss(outs[0], outs[1], outs[2], c["MARGIN_1"], c["LAM"], distances, d_max_squared,
                                         f_max_squared, trip, dl)
        if name in ("distance_quadruplet", "distance_lazy_quadruplet", "huber_distance_quadruplet",
                    "huber_distance_lazy_quadruplet"):                        # train.py:729-763
            trip = "lazy_triplet_loss" if "lazy" in name else "triplet_loss"
            dl = "huber_distance_loss" if "huber" in name else "distance_loss"
            return distance_quadruplet_loss(outs[0], outs[1], outs[2], outs[3], c["MARGIN_1"], c["MARGIN_2"], c["LAM"],
                                            distances, d_max_squared, f_max_squared, trip, dl)
        if name == "ms_loss":                                                 # train.py:821-827
            return ms_loss(ms_labels(T, P, N), output, ms_mining=c["MSMINING"])
        if name == "wms":                                                     # train.py:851-852
            return wms_loss(distances, output, d_alpha=c["ALPHA"], d_beta=c["BETA"], wfunction=c["WFUNCTION"],
                            sumfunction=c["SUMFUNCTION"])
        if name == "logratio":                                                # train.py:687-691, 854-855
            d = _f32(distances).reshape(T, P + N)
            return logratio_loss(outs[0], outs[1], outs[2], d[:, :P].reshape(T, P, 1), d[:, P:].reshape(T, N, 1))
        raise AssertionError(name)

    f.__name__ = f"loss_{name}"
    return f
