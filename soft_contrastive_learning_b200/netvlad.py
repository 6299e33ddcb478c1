"""NetVLAD aggregation head and PCA-whitening projection, by the reference's names.

Host-side mirror of
  x = tf.nn.l2_normalize(x, axis=-1); x = layers.netVLAD(x, 64)      /root/reference/model/nets.py:66-67
  x_tf = tf.matmul(full_out - m, v, adjoint_b=True); output = x_tf / tf.sqrt(var)   train/train.py:646-652
Forward and backward run in libscl_b200.so (csrc/netvlad.cu); torch owns memory, streams and the autograd tape.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib
from .losses import _f32, _p, _stream, _ws


class _NetVladFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, assign_w, centers):
        B = x.shape[0]
        Cc = x.shape[-1]
        K = assign_w.shape[-1]
        x3 = _f32(x.detach()).reshape(B, -1, Cc)
        w = _f32(assign_w.detach()).reshape(Cc, K)
        c = _f32(centers.detach()).reshape(Cc, K)
        HW = x3.shape[1]
        L = lib()
        nbytes = C.c_size_t()
        check(L.scl_netvlad_workspace_bytes(B, HW, Cc, K, C.byref(nbytes)), "scl_netvlad_workspace_bytes")
        ws = _ws(nbytes.value, x3.device)
        out = torch.empty((B, Cc * K), dtype=torch.float32, device=x3.device)
        check(L.scl_netvlad_fwd(_p(x3), _p(w), _p(c), B, HW, Cc, K, _p(out), _p(ws), ws.numel(), _stream()),
              "scl_netvlad_fwd")
        ctx.save_for_backward(x3, w, c, ws)
        ctx.dims = (B, HW, Cc, K, x.shape, assign_w.shape, centers.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        x3, w, c, ws = ctx.saved_tensors
        B, HW, Cc, K, xs, ws_shape, cs = ctx.dims
        dout = _f32(dout)
        need_x, need_w, need_c = ctx.needs_input_grad
        dx = torch.empty_like(x3) if need_x else None
        dw = torch.empty_like(w) if need_w else None
        dc = torch.empty_like(c) if need_c else None
        check(lib().scl_netvlad_bwd(_p(x3), _p(w), _p(c), _p(dout), B, HW, Cc, K, _p(dx), _p(dw), _p(dc), _p(ws),
                                    ws.numel(), _stream()), "scl_netvlad_bwd")
        return (None if dx is None else dx.reshape(xs), None if dw is None else dw.reshape(ws_shape),
                None if dc is None else dc.reshape(cs))


def netVLAD(inputs, assignment_kernel, cluster_centers, num_clusters=64):
    """``layers.netVLAD(tf.nn.l2_normalize(inputs, -1), num_clusters)`` (model/nets.py:66-67).

    inputs [B,h,w,C] (NHWC conv5_3 map) or [B,HW,C]; assignment_kernel [1,1,C,K] or [C,K] ('assignment/kernel');
    cluster_centers [1,1,1,C,K] or [C,K] ('cluster_centers').  TF variables become explicit tensors.
    Returns [B, C*K] (index c*K+k), intra-normalised and L2-normalised."""
    if assignment_kernel.shape[-1] != num_clusters:
        raise ValueError("assignment kernel does not match num_clusters")
    t = [a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))
         for a in (inputs, assignment_kernel, cluster_centers)]
    dev = torch.device("cuda", torch.cuda.current_device())
    t = [a if a.is_cuda else a.to(dev) for a in t]
    out = _NetVladFn.apply(*t)
    return out.cpu().numpy() if isinstance(inputs, np.ndarray) else out


def _pca_ws(B, Din, Dout, dev):
    n = C.c_size_t()
    check(lib().scl_pca_workspace_bytes(B, Din, Dout, C.byref(n)), "scl_pca_workspace_bytes")
    return _ws(n.value, dev)


def set_gemm_precision(mode):
    """0: fp32-grade 3xTF32 tensor-core contractions (default); 1: single TF32 pass (stated tolerance 2e-3)."""
    check(lib().scl_set_gemm_precision(int(mode)), "scl_set_gemm_precision")


class PreparedPCA:
    """The projection matrix split once into fp16 hi / lo halves on the device (``scl_pca_prepare``); projections then run
    on the pre-split f16 tensor-core engine.  ``v`` is a fed constant in the reference (train/train.py:281-283), so
    :func:`pca_project` builds this lazily and re-uses it while the tensor it was built from is unchanged."""

    def __init__(self, v):
        v2 = _f32(v.detach())
        self.Dout, self.Din = v2.shape
        n = C.c_size_t()
        check(lib().scl_pca_shadow_bytes(self.Din, self.Dout, C.byref(n)), "scl_pca_shadow_bytes")
        self.shadow = _ws(n.value, v2.device)
        check(lib().scl_pca_prepare(_p(v2), self.Din, self.Dout, _p(self.shadow), self.shadow.numel(), _stream()),
              "scl_pca_prepare")
        self.key = _pca_key(v)

    @staticmethod
    def supported(Din, Dout):
        return Din % 8 == 0 and Dout % 8 == 0 and 8 <= Din <= 32768 and Dout <= 32768


def _pca_key(v):
    return (v.data_ptr(), v._version, tuple(v.shape), str(v.device))


_prepared = {}          # id(tensor) -> (weakref to the tensor, PreparedPCA)


def _prepared_for(v):
    """The PreparedPCA of a LIVE device tensor (same object, same version counter), built on first use.  Host arrays and
    temporaries are not cached: a recycled device address must never resurrect another matrix's shadow."""
    import weakref
    if not (isinstance(v, torch.Tensor) and v.is_cuda and v.dtype == torch.float32 and v.dim() == 2 and v.is_contiguous()
            and PreparedPCA.supported(v.shape[1], v.shape[0])):
        return None
    ent = _prepared.get(id(v))
    if ent is not None and ent[0]() is v and ent[1].key == _pca_key(v):
        return ent[1]
    for k in [k for k, e in _prepared.items() if e[0]() is None]:
        del _prepared[k]
    p = PreparedPCA(v)
    _prepared[id(v)] = (weakref.ref(v), p)
    return p


class _PcaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, v, m, var):
        x2, v2, m1, var1 = _f32(x.detach()), _f32(v.detach()), _f32(m.detach()), _f32(var.detach())
        B, Din = x2.shape
        Dout = v2.shape[0]
        y = torch.empty((B, Dout), dtype=torch.float32, device=x2.device)
        prep = _prepared_for(v) if lib().scl_get_gemm_precision() == 0 else None
        if prep is not None:
            n = C.c_size_t()
            check(lib().scl_pca_prepared_workspace_bytes(B, Din, Dout, C.byref(n)), "scl_pca_prepared_workspace_bytes")
            ws = _ws(n.value, x2.device)
            check(lib().scl_pca_fwd_prepared(_p(x2), _p(prep.shadow), _p(m1), _p(var1), B, Din, Dout, _p(y), _p(ws),
                                             ws.numel(), _stream()), "scl_pca_fwd_prepared")
        else:
            ws = _pca_ws(B, Din, Dout, x2.device)
            check(lib().scl_pca_fwd(_p(x2), _p(v2), _p(m1), _p(var1), B, Din, Dout, _p(y), _p(ws), ws.numel(), _stream()),
                  "scl_pca_fwd")
        ctx.save_for_backward(v2, var1)
        ctx.prep = prep
        ctx.dims = (B, Din, Dout)
        return y

    @staticmethod
    def backward(ctx, dy):
        v2, var1 = ctx.saved_tensors
        B, Din, Dout = ctx.dims
        dy = _f32(dy)
        dx = torch.empty((B, Din), dtype=torch.float32, device=dy.device)
        if ctx.prep is not None:
            n = C.c_size_t()
            check(lib().scl_pca_prepared_workspace_bytes(B, Din, Dout, C.byref(n)), "scl_pca_prepared_workspace_bytes")
            ws = _ws(n.value, dy.device)
            check(lib().scl_pca_bwd_prepared(_p(dy), _p(ctx.prep.shadow), _p(var1), B, Din, Dout, _p(dx), _p(ws), ws.numel(),
                                             _stream()), "scl_pca_bwd_prepared")
        else:
            ws = _pca_ws(B, Din, Dout, dy.device)
            check(lib().scl_pca_bwd(_p(dy), _p(v2), _p(var1), B, Din, Dout, _p(dx), _p(ws), ws.numel(), _stream()),
                  "scl_pca_bwd")
        return dx, None, None, None          # v, m, var are fed placeholders, not trained (train.py:647-649)


def pca_project(full_out, v, m, var):
    """train/train.py:650-651: ``matmul(full_out - m, v, adjoint_b=True) / sqrt(var)``.
    full_out [B,Din], v [Dout,Din], m [Din], var [Dout] -> [B,Dout]."""
    t = [a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a)) for a in (full_out, v, m, var)]
    dev = torch.device("cuda", torch.cuda.current_device())
    t = [a if a.is_cuda else a.to(dev) for a in t]
    out = _PcaFn.apply(*t)
    return out.cpu().numpy() if isinstance(full_out, np.ndarray) else out


def _gemm(A, B, M, N, K, a_mn, b_mn, precision=0):
    """C[M,N] = A . B^T on the tcgen05 TF32 engine (csrc/tc_gemm.cu); leading dimensions from the tensors' strides."""
    ldc = (N + 3) // 4 * 4
    out = torch.empty((M, ldc), dtype=torch.float32, device=A.device)
    check(lib().scl_gemm_tf32(_p(A), _p(B), _p(out), M, N, K, A.stride(0), B.stride(0), ldc, int(a_mn), int(b_mn), None,
                              int(precision), _stream()), "scl_gemm_tf32")
    return out[:, :N]


def _pad_cols(t, mult=4):
    """A copy of the 2-D tensor whose row pitch is a multiple of `mult` floats (zero padding), viewed at its own width."""
    n, w = t.shape
    wp = (w + mult - 1) // mult * mult
    if wp == w and t.is_contiguous():
        return t
    buf = torch.zeros((n, wp), dtype=torch.float32, device=t.device)
    buf[:, :w] = t
    return buf[:, :w]


def pca_fit(features, n_components, return_device=False):
    """``PCA(whiten=True, n_components=d).fit(pca_f)`` of /root/reference/evaluation/top-n.py:74-75 on the GPU
    (SURVEY 8f row 4), as the exact truncated decomposition (scikit-learn's ``svd_solver='full'`` result; the
    reference's default ``'auto'`` resolves to the randomized solver with an unseeded RNG for these sizes, i.e. to an
    approximation of what is computed here).  Returns ``(v, m, var)`` = (components_ [d,D], mean_ [D],
    explained_variance_ [d]) ready for :func:`pca_project`.

    mean and centring: ``scl_pca_center``; n <= D: Gram ``G = Xc Xc^T`` [n,n] and the back-projection
    ``V = (U / sigma)^T Xc`` on the fp32-grade tcgen05 GEMM, with ``G = U diag(sigma^2) U^T``; n > D: covariance
    ``Xc^T Xc`` [D,D] on the same GEMM.  Only the small symmetric eigenproblem (n x n or D x D, float64) is a
    library call (``torch.linalg.eigh``).  Signs follow scikit-learn's ``svd_flip(u_based_decision=False)``: the
    largest-magnitude entry of every component is positive.  ``explained_variance_ = sigma^2 / (n - 1)``."""
    x = features if isinstance(features, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(features, dtype=np.float32))
    dev = torch.device("cuda", torch.cuda.current_device())
    x = _f32(x if x.is_cuda else x.to(dev))
    n, D = x.shape
    d = int(n_components)
    if not 1 <= d <= min(n, D):
        raise ValueError(f"n_components={d} must be between 1 and min(n_samples, n_features)={min(n, D)}")
    nbytes = C.c_size_t()
    check(lib().scl_pca_center_workspace_bytes(n, D, C.byref(nbytes)), "scl_pca_center_workspace_bytes")
    ws = _ws(nbytes.value, x.device)
    mean = torch.empty(D, dtype=torch.float32, device=x.device)
    Dp = (D + 3) // 4 * 4
    xc_buf = torch.zeros((n, Dp), dtype=torch.float32, device=x.device) if Dp != D else torch.empty((n, D), dtype=torch.float32, device=x.device)
    if Dp == D:
        check(lib().scl_pca_center(_p(x), n, D, _p(mean), _p(xc_buf), _p(ws), ws.numel(), _stream()), "scl_pca_center")
        xc = xc_buf
    else:                                         # ragged D: centre densely, then re-pitch for the TMA path
        dense = torch.empty((n, D), dtype=torch.float32, device=x.device)
        check(lib().scl_pca_center(_p(x), n, D, _p(mean), _p(dense), _p(ws), ws.numel(), _stream()), "scl_pca_center")
        xc_buf[:, :D] = dense
        xc = xc_buf[:, :D]
    if n <= D:
        G = _gemm(xc, xc, n, n, D, False, False)                           # Xc Xc^T
        lam, U = torch.linalg.eigh(G.double())                             # ascending
        lam = lam.flip(0)[:d].clamp_min(0.0)
        U = U.flip(1)[:, :d]
        sigma = lam.sqrt()
        Us = _pad_cols((U / sigma.clamp_min(1e-300)).float())              # [n,d]
        comps = _gemm(Us, xc, d, D, n, True, True)                         # A(m,k) = Us[k,m], B(j,k) = Xc[k,j]
    else:
        Cv = _gemm(xc, xc, D, D, n, True, True)                            # Xc^T Xc
        lam, V = torch.linalg.eigh(Cv.double())
        lam = lam.flip(0)[:d].clamp_min(0.0)
        comps = V.flip(1)[:, :d].t().float()
    comps = comps.contiguous()
    idx = comps.abs().argmax(dim=1, keepdim=True)
    comps = comps * torch.sign(torch.gather(comps, 1, idx))
    var = (lam / max(n - 1, 1)).float()
    if return_device:
        return comps, mean, var
    return comps.cpu().numpy(), mean.cpu().numpy(), var.cpu().numpy()


def pca_from_sklearn(pca):
    """(v, m, var) of a fitted ``sklearn.decomposition.PCA(whiten=True)`` (evaluation/top-n.py:74-75):
    pca.transform(x) == pca_project(x, components_, mean_, explained_variance_)."""
    return (np.asarray(pca.components_, dtype=np.float32), np.asarray(pca.mean_, dtype=np.float32),
            np.asarray(pca.explained_variance_, dtype=np.float32))
