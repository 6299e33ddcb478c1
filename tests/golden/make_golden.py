"""Generate tests/golden/*.npz by running the REFERENCE'S OWN model/losses.py over the NumPy TF shim.

Run in the build container (needs /root/reference):   python tests/golden/make_golden.py
The GPU box has no /root/reference; tests there read only the committed .npz files.

For every case the file stores the float32 inputs, the reference forward value (float64, produced by the
reference source), the oracle's value (must agree to 1e-12) and the oracle's float64 autograd gradient,
which is itself checked here against central finite differences OF THE REFERENCE FORWARD.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import torch  # noqa: E402

from oracle import losses as ol  # noqa: E402
from soft_contrastive_learning_b200 import synth  # noqa: E402
from tf_numpy_shim import A, load_reference_losses  # noqa: E402

ref, pn = load_reference_losses()


def fd_grad(f, x, idxs, h=1e-6):
    """Central differences of scalar f at flat positions idxs of x."""
    g = np.zeros(len(idxs))
    flat = x.reshape(-1)
    for n, i in enumerate(idxs):
        old = flat[i]
        flat[i] = old + h
        fp = float(f(x))
        flat[i] = old - h
        fm = float(f(x))
        flat[i] = old
        g[n] = (fp - fm) / (2 * h)
    return g


def check_grad(name, f_ref, x64, g_oracle, n=24, seed=0, tol=2e-5):
    rng = np.random.default_rng(seed)
    idxs = rng.choice(x64.size, size=min(n, x64.size), replace=False)
    g_fd = fd_grad(f_ref, x64.copy(), idxs)
    g_or = g_oracle.reshape(-1)[idxs]
    scale = max(np.abs(g_oracle).max(), 1e-12)
    err = np.abs(g_fd - g_or).max() / scale
    assert err < tol, (name, err)
    return err


def save(name, **kw):
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **kw)
    print("wrote", name, {k: (np.asarray(v).shape if hasattr(v, "shape") else v) for k, v in kw.items()})


def make_wms():
    rng = np.random.default_rng(7)
    P = N = 12
    S = 1 + P + N
    D = 64
    xy = synth.tuple_xy(rng, 1, P, N)
    emb = synth.tuple_descriptors(rng, 1, P, N, D)[0]                 # [S,D] float32
    dist = synth.pairwise_euclid(xy)[0].astype(np.float32)            # [S,S]
    out = {"emb": emb, "dist": dist}
    e64, d64 = emb.astype(np.float64), dist.astype(np.float64)
    variants = [
        ("exp_ms_mine", dict(wfunction="exp", sumfunction="ms", ms_mining=True)),
        ("exp_ms_nomine", dict(wfunction="exp", sumfunction="ms", ms_mining=False)),
        ("lin_ms_mine", dict(wfunction="lin", sumfunction="ms", ms_mining=True)),
        ("tanh_ms_mine", dict(wfunction="tanh", sumfunction="ms", ms_mining=True)),
        ("exp_plain_mine", dict(wfunction="exp", sumfunction="plain", ms_mining=True)),
    ]
    for tag, kw in variants:
        f2 = lambda e, kw=kw: ref.wms_loss(A(dist, np.float32), A(e), 0.8, 15.0, **kw)
        v_ref = float(f2(e64))
        # the call train.py:852 actually makes: distances [1,S,S], output [S,D]
        v_ref3 = float(ref.wms_loss(A(dist[None], np.float32), A(e64), 0.8, 15.0, **kw))
        v_or, (g_or,) = ol.value_and_grad(
            lambda e: ol.wms_loss(torch.as_tensor(d64), e, 0.8, 15.0, **kw), [e64])
        assert abs(v_ref - v_or) < 1e-6 * max(1, abs(v_ref)), (tag, v_ref, v_or)   # float32 exp: NumPy vs torch differ by an ulp
        assert abs(v_ref - v_ref3) < 1e-9, (tag, v_ref, v_ref3)
        err = check_grad("wms_" + tag, f2, e64, g_or)
        _, mp, mn = ol.wms_loss(d64, e64, 0.8, 15.0, return_masks=True, **kw)
        out["loss_" + tag] = v_ref
        out["grad_" + tag] = g_or
        out["keptpos_" + tag] = mp.numpy()
        out["keptneg_" + tag] = mn.numpy()
        print(f"  wms {tag}: loss={v_ref:.10f} fd_err={err:.2e} kept_pos={int(mp.sum())} kept_neg={int(mn.sum())}")
    save("wms_flat_S25_D64", **out)

    # tuple mode T=4 (oracle-defined batching: mean over tuples of the per-tuple reference loss)
    rng = np.random.default_rng(11)
    T, D = 4, 256
    xy = synth.tuple_xy(rng, T, P, N)
    emb = synth.tuple_descriptors(rng, T, P, N, D)
    dist = synth.pairwise_euclid(xy).astype(np.float32)
    per = [float(ref.wms_loss(A(dist[t], np.float32), A(emb[t].astype(np.float64)), 0.8, 15.0)) for t in range(T)]
    v_or, (g_or,) = ol.value_and_grad(
        lambda e: ol.wms_loss_tuples(torch.as_tensor(dist.astype(np.float64)), e, 0.8, 15.0), [emb.astype(np.float64)])
    assert abs(np.mean(per) - v_or) < 1e-6
    save("wms_tuples_T4_S25_D256", emb=emb, dist=dist, loss=np.mean(per), per_tuple=np.array(per), grad=g_or)


def make_ms():
    rng = np.random.default_rng(3)
    T, P, N, D = 3, 4, 5, 48
    emb = synth.tuple_descriptors(rng, T, P, N, D).reshape(T * (1 + P + N), D)
    labels = ol.ms_labels(T, P, N)
    e64 = emb.astype(np.float64)
    out = {"emb": emb, "labels": labels}
    for tag, mining in (("mine", True), ("nomine", False)):
        f = lambda e, mining=mining: ref.ms_loss(A(labels), A(e), ms_mining=mining)
        v_ref = float(f(e64))
        v_or, (g_or,) = ol.value_and_grad(lambda e: ol.ms_loss(labels, e, ms_mining=mining), [e64])
        assert abs(v_ref - v_or) < 1e-12 * max(1, abs(v_ref)), (tag, v_ref, v_or)
        err = check_grad("ms_" + tag, f, e64, g_or)
        out["loss_" + tag] = v_ref
        out["grad_" + tag] = g_or
        print(f"  ms {tag}: loss={v_ref:.10f} fd_err={err:.2e}")
    save("ms_T3_P4_N5_D48", **out)


def make_tuple_losses():
    rng = np.random.default_rng(5)
    T, P, N, D = 3, 4, 6, 40
    xy = synth.tuple_xy(rng, T, P, N, other=True)
    emb = (0.1 * synth.tuple_descriptors(rng, T, P, N, D, other=True, pos_noise=1.2)).astype(np.float32)  # hinges partly active
    e64 = emb.astype(np.float64)
    sqd = synth.anchor_sq_dists(xy, P).astype(np.float32)
    m1, m2, lam = 0.1, 0.2, 0.5
    dmax, fmax = 225.0, 2.0
    out = {"emb": emb, "sq_d_dists": sqd, "P": P, "N": N, "m1": m1, "m2": m2, "lam": lam,
           "d_max_squared": dmax, "f_max_squared": fmax}

    def sp(e):
        return [A(x) for x in ol.split_tuple(e, P, N, other=True)]

    def spt(e):
        return ol.split_tuple(e, P, N, other=True)

    sq64 = sqd.astype(np.float64)
    cases = {
        "triplet": (lambda e: pn.triplet_loss(*sp(e)[:3], m1), lambda e: ol.triplet_loss(*spt(e)[:3], m1)),
        "lazy_triplet": (lambda e: pn.lazy_triplet_loss(*sp(e)[:3], m1), lambda e: ol.lazy_triplet_loss(*spt(e)[:3], m1)),
        "quadruplet": (lambda e: pn.quadruplet_loss(*sp(e), m1, m2), lambda e: ol.quadruplet_loss(*spt(e), m1, m2)),
        "lazy_quadruplet": (lambda e: pn.lazy_quadruplet_loss(*sp(e), m1, m2),
                            lambda e: ol.lazy_quadruplet_loss(*spt(e), m1, m2)),
        # in-repo twins, reference source itself (losses.py:63-73, 197-214)
        "evil_triplet": (lambda e: ref.evil_triplet_loss(*sp(e)[:3], m1), lambda e: ol.evil_triplet_loss(*spt(e)[:3], m1)),
        "evil_quadruplet": (lambda e: ref.evil_quadruplet_loss(*sp(e), m1, m2),
                            lambda e: ol.evil_quadruplet_loss(*spt(e), m1, m2)),
        # reference source (losses.py:239-264) dispatching by name into the pointnetvlad shim
        "huber_distance_triplet": (
            lambda e: ref.distance_triplet_loss(*sp(e)[:3], m1, lam, A(sq64), dmax, fmax, "triplet_loss", "huber_distance_loss"),
            lambda e: ol.distance_triplet_loss(*spt(e)[:3], m1, lam, torch.as_tensor(sq64), dmax, fmax,
                                               "triplet_loss", "huber_distance_loss")),
        "huber_distance_lazy_triplet": (
            lambda e: ref.distance_triplet_loss(*sp(e)[:3], m1, lam, A(sq64), dmax, fmax, "lazy_triplet_loss", "huber_distance_loss"),
            lambda e: ol.distance_triplet_loss(*spt(e)[:3], m1, lam, torch.as_tensor(sq64), dmax, fmax,
                                               "lazy_triplet_loss", "huber_distance_loss")),
        "distance_triplet": (
            lambda e: ref.distance_triplet_loss(*sp(e)[:3], m1, lam, A(sq64), dmax, fmax, "triplet_loss", "distance_loss"),
            lambda e: ol.distance_triplet_loss(*spt(e)[:3], m1, lam, torch.as_tensor(sq64), dmax, fmax,
                                               "triplet_loss", "distance_loss")),
    }
    # distance_quadruplet_loss, reference source (losses.py:267-307; dispatch train.py:729-763)
    for tname, tl in (("", "triplet_loss"), ("lazy_", "lazy_triplet_loss")):
        for dname, dl in (("distance", "distance_loss"), ("huber_distance", "huber_distance_loss")):
            cases[f"{dname}_{tname}quadruplet"] = (
                lambda e, tl=tl, dl=dl: ref.distance_quadruplet_loss(*sp(e), m1, m2, lam, A(sq64), dmax, fmax, tl, dl),
                lambda e, tl=tl, dl=dl: ol.distance_quadruplet_loss(*spt(e), m1, m2, lam, torch.as_tensor(sq64), dmax,
                                                                    fmax, tl, dl))
    for tag, (f_ref, f_or) in cases.items():
        v_ref = float(f_ref(e64))
        v_or, (g_or,) = ol.value_and_grad(f_or, [e64])
        assert abs(v_ref - v_or) < 1e-12 * max(1, abs(v_ref)), (tag, v_ref, v_or)
        err = check_grad(tag, f_ref, e64, g_or)
        out["loss_" + tag] = v_ref
        out["grad_" + tag] = g_or
        print(f"  {tag}: loss={v_ref:.10f} fd_err={err:.2e} nz_grad={int((g_or != 0).sum())}")
    save("tuple_losses_T3_P4_N6_D40", **out)


def make_pairwise_distance_loss():
    """pairwise_distance_loss (losses.py:627-646) on [anchor, positives] with all-pairs squared metres."""
    rng = np.random.default_rng(8)
    T, P, D = 3, 5, 40
    xy = synth.tuple_xy(rng, T, P, 2)[:, :P + 1]
    emb = (0.6 * synth.tuple_descriptors(rng, T, P, 2, D, pos_noise=1.2)[:, :P + 1]).astype(np.float32)
    e64 = emb.astype(np.float64)
    diff = xy[:, :, None, :] - xy[:, None, :, :]
    sqd = (diff ** 2).sum(-1).astype(np.float32)                     # 'pairwise' DISTANCE_TYPE, train.py:535-537
    sq64 = sqd.astype(np.float64)
    dmax, fmax = 225.0, 2.0
    out = {"emb": emb, "pairwise_sq_d": sqd, "P": P, "d_max_squared": dmax, "f_max_squared": fmax}
    for tag, dl in (("squared", "distance_loss"), ("huber", "huber_distance_loss")):
        f_ref = lambda e, dl=dl: ref.pairwise_distance_loss(A(e[:, :1]), A(e[:, 1:]), A(sq64), dmax, fmax, dl)
        f_or = lambda e, dl=dl: ol.pairwise_distance_loss(e[:, :1], e[:, 1:], torch.as_tensor(sq64), dmax, fmax, dl)
        v_ref = float(f_ref(e64))
        v_or, (g_or,) = ol.value_and_grad(f_or, [e64])
        assert abs(v_ref - v_or) < 1e-12 * max(1, abs(v_ref)), (tag, v_ref, v_or)
        err = check_grad("pairwise_" + tag, f_ref, e64, g_or)
        out["loss_" + tag] = v_ref
        out["grad_" + tag] = g_or
        print(f"  pairwise_distance_loss[{tag}]: loss={v_ref:.10f} fd_err={err:.2e}")
    save("pairwise_distance_loss_T3_P5_D40", **out)


def make_logratio():
    rng = np.random.default_rng(9)
    P = N = 5
    D = 32
    xy = synth.tuple_xy(rng, 1, P, N)
    emb = synth.tuple_descriptors(rng, 1, P, N, D)
    e64 = emb.astype(np.float64)
    sp_, sn_ = synth.logratio_sq_dists(xy, P, N)
    sp_, sn_ = sp_.astype(np.float32), sn_.astype(np.float32)

    def f_ref(e):
        a, p, n = ol.split_tuple(e, P, N)
        return ref.logratio_loss(A(a), A(p), A(n), A(sp_.astype(np.float64)), A(sn_.astype(np.float64)))

    def f_or(e):
        a, p, n = ol.split_tuple(e, P, N)
        return ol.logratio_loss(a, p, n, torch.as_tensor(sp_.astype(np.float64)), torch.as_tensor(sn_.astype(np.float64)))

    v_ref = float(f_ref(e64))
    v_or, (g_or,) = ol.value_and_grad(f_or, [e64])
    assert abs(v_ref - v_or) < 1e-12 * max(1, abs(v_ref)), (v_ref, v_or)
    err = check_grad("logratio", f_ref, e64, g_or)
    print(f"  logratio: loss={v_ref:.10f} fd_err={err:.2e}")
    save("logratio_P5_N5_D32", emb=emb, sq_pos=sp_, sq_neg=sn_, loss=v_ref, grad=g_or)


def make_pairwise():
    # the tensor hard-coded in the reference's only self-check, model/losses.py:708-710
    B = np.array([[[1.0, 1], [2, 2], [3, 3]], [[1, 1], [2, 2], [4, 4]]])
    d_ref = np.asarray(ref._pairwise_squared_distances(A(B)))
    d_or = ol.pairwise_squared_distances(B).numpy()
    expect = np.array([[[0, 2, 8], [2, 0, 2], [8, 2, 0]], [[0, 2, 18], [2, 0, 8], [18, 8, 0]]], dtype=np.float64)
    assert np.array_equal(d_ref, expect) and np.array_equal(d_or, expect)
    rng = np.random.default_rng(2)
    X = rng.standard_normal((3, 7, 33)).astype(np.float32)
    d_ref2 = np.asarray(ref._pairwise_squared_distances(A(X.astype(np.float64))))
    assert np.allclose(d_ref2, ol.pairwise_squared_distances(X.astype(np.float64)).numpy(), rtol=0, atol=1e-12)
    save("pairwise_sqdist", selfcheck_in=B, selfcheck_out=expect, x=X, d=d_ref2)


if __name__ == "__main__":
    make_wms()
    make_ms()
    make_tuple_losses()
    make_logratio()
    make_pairwise()
    make_pairwise_distance_loss()
