"""CPU: the oracle reproduces the golden vectors that the REFERENCE SOURCE produced over the NumPy TF shim."""
import numpy as np
import pytest
import torch

from oracle import losses as ol
from oracle import netvlad as onv
from oracle import retrieval as orr
from soft_contrastive_learning_b200 import synth

WMS_VARIANTS = {
    "exp_ms_mine": dict(wfunction="exp", sumfunction="ms", ms_mining=True),
    "exp_ms_nomine": dict(wfunction="exp", sumfunction="ms", ms_mining=False),
    "lin_ms_mine": dict(wfunction="lin", sumfunction="ms", ms_mining=True),
    "tanh_ms_mine": dict(wfunction="tanh", sumfunction="ms", ms_mining=True),
    "exp_plain_mine": dict(wfunction="exp", sumfunction="plain", ms_mining=True),
}


@pytest.mark.parametrize("tag", list(WMS_VARIANTS))
def test_wms_flat_golden(golden, tag):
    g = golden("wms_flat_S25_D64")
    e, d = g["emb"].astype(np.float64), g["dist"].astype(np.float64)
    v, (gr,) = ol.value_and_grad(lambda x: ol.wms_loss(torch.as_tensor(d), x, 0.8, 15.0, **WMS_VARIANTS[tag]), [e])
    # the golden value came out of the reference source over NumPy; its float32 exp differs from torch's by an ulp
    assert abs(v - float(g["loss_" + tag])) <= 1e-6 * max(1.0, abs(v))
    assert np.allclose(gr, g["grad_" + tag], rtol=1e-10, atol=1e-14)
    _, mp, mn = ol.wms_loss(d, e, 0.8, 15.0, return_masks=True, **WMS_VARIANTS[tag])
    assert np.array_equal(mp.numpy(), g["keptpos_" + tag]) and np.array_equal(mn.numpy(), g["keptneg_" + tag])


def test_wms_tuple_mode_is_mean_of_reference_tuples(golden):
    g = golden("wms_tuples_T4_S25_D256")
    e, d = g["emb"].astype(np.float64), g["dist"].astype(np.float64)
    v = float(ol.wms_loss_tuples(torch.as_tensor(d), torch.as_tensor(e), 0.8, 15.0))
    assert abs(v - float(g["loss"])) < 1e-6
    assert abs(v - g["per_tuple"].mean()) < 1e-6
    # T=1 tuple mode == the flat 2-D call
    v1 = float(ol.wms_loss_tuples(torch.as_tensor(d[:1]), torch.as_tensor(e[:1]), 0.8, 15.0))
    assert abs(v1 - g["per_tuple"][0]) < 1e-6


@pytest.mark.parametrize("tag,mining", [("mine", True), ("nomine", False)])
def test_ms_golden(golden, tag, mining):
    g = golden("ms_T3_P4_N5_D48")
    v, (gr,) = ol.value_and_grad(lambda x: ol.ms_loss(g["labels"], x, ms_mining=mining), [g["emb"].astype(np.float64)])
    assert abs(v - float(g["loss_" + tag])) <= 1e-12
    assert np.allclose(gr, g["grad_" + tag], rtol=1e-10, atol=1e-14)


def test_ms_labels_match_train_py():
    lab = ol.ms_labels(2, 2, 3)
    assert lab.tolist() == [0, 0, 0, 1, 2, 3, 4, 4, 4, 5, 6, 7]


TUPLE_CASES = ["triplet", "lazy_triplet", "quadruplet", "lazy_quadruplet", "evil_triplet", "evil_quadruplet",
               "huber_distance_triplet", "huber_distance_lazy_triplet", "distance_triplet"]


def _tuple_oracle(tag, g, e):
    P, N = int(g["P"]), int(g["N"])
    m1, m2, lam = float(g["m1"]), float(g["m2"]), float(g["lam"])
    sq = torch.as_tensor(g["sq_d_dists"].astype(np.float64))
    q, p, n, o = ol.split_tuple(e, P, N, other=True)
    if tag == "triplet":
        return ol.triplet_loss(q, p, n, m1)
    if tag == "lazy_triplet":
        return ol.lazy_triplet_loss(q, p, n, m1)
    if tag == "quadruplet":
        return ol.quadruplet_loss(q, p, n, o, m1, m2)
    if tag == "lazy_quadruplet":
        return ol.lazy_quadruplet_loss(q, p, n, o, m1, m2)
    if tag == "evil_triplet":
        return ol.evil_triplet_loss(q, p, n, m1)
    if tag == "evil_quadruplet":
        return ol.evil_quadruplet_loss(q, p, n, o, m1, m2)
    trip = "lazy_triplet_loss" if "lazy" in tag else "triplet_loss"
    dl = "huber_distance_loss" if "huber" in tag else "distance_loss"
    return ol.distance_triplet_loss(q, p, n, m1, lam, sq, float(g["d_max_squared"]), float(g["f_max_squared"]), trip, dl)


@pytest.mark.parametrize("tag", TUPLE_CASES)
def test_tuple_losses_golden(golden, tag):
    g = golden("tuple_losses_T3_P4_N6_D40")
    v, (gr,) = ol.value_and_grad(lambda x: _tuple_oracle(tag, g, x), [g["emb"].astype(np.float64)])
    assert abs(v - float(g["loss_" + tag])) <= 1e-12
    assert np.allclose(gr, g["grad_" + tag], rtol=1e-10, atol=1e-14)
    assert float(g["loss_" + tag]) > 0.0          # the fixture exercises active hinges


def test_logratio_golden(golden):
    g = golden("logratio_P5_N5_D32")
    P = N = 5

    def f(x):
        a, p, n = ol.split_tuple(x, P, N)
        return ol.logratio_loss(a, p, n, torch.as_tensor(g["sq_pos"].astype(np.float64)),
                                torch.as_tensor(g["sq_neg"].astype(np.float64)))
    v, (gr,) = ol.value_and_grad(f, [g["emb"].astype(np.float64)])
    assert abs(v - float(g["loss"])) <= 1e-12 * abs(v)
    assert np.allclose(gr, g["grad"], rtol=1e-10, atol=1e-14)


def test_pairwise_selfcheck_of_reference(golden):
    g = golden("pairwise_sqdist")
    assert np.array_equal(ol.pairwise_squared_distances(g["selfcheck_in"]).numpy(), g["selfcheck_out"])
    assert np.allclose(ol.pairwise_squared_distances(g["x"].astype(np.float64)).numpy(), g["d"], atol=1e-12)


def test_oracle_gradient_finite_differences():
    rng = np.random.default_rng(0)
    e = rng.standard_normal((6, 8))
    xy = rng.uniform(0, 40, size=(6, 2))
    d = np.sqrt(((xy[:, None] - xy[None]) ** 2).sum(-1))
    f = lambda x: ol.wms_loss(torch.as_tensor(d), x, 0.8, 15.0)
    v, (g,) = ol.value_and_grad(f, [e])
    h = 1e-6
    for (i, j) in [(0, 0), (2, 3), (5, 7)]:
        ep, em = e.copy(), e.copy()
        ep[i, j] += h
        em[i, j] -= h
        fd = (float(f(torch.as_tensor(ep))) - float(f(torch.as_tensor(em)))) / (2 * h)
        assert abs(fd - g[i, j]) < 1e-6 * max(1.0, abs(fd))


def test_netvlad_oracle_matches_5d_formulation():
    """The oracle's einsum form equals the upstream graph's explicit [B,HW,D,K] residual tensor."""
    rng = np.random.default_rng(1)
    B, HW, D, K = 2, 7, 16, 4
    x = torch.as_tensor(rng.standard_normal((B, HW, D)))
    w = torch.as_tensor(0.3 * rng.standard_normal((D, K)))
    c = torch.as_tensor(0.3 * rng.standard_normal((D, K)))
    out = onv.netvlad_head(x, w, c)
    xh = x / torch.sqrt(torch.clamp((x * x).sum(-1, keepdim=True), min=1e-12))
    a = torch.softmax(xh @ w, dim=-1)
    v = (a[:, :, None, :] * (xh[:, :, :, None] + c[None, None])).sum(1)          # [B,D,K]
    v = v.permute(0, 2, 1)
    v = v / torch.sqrt((v ** 2).sum(-1, keepdim=True) + 1e-12)
    v = v.permute(0, 2, 1).reshape(B, -1)
    v = v / torch.sqrt((v ** 2).sum(-1, keepdim=True) + 1e-12)
    assert torch.allclose(out, v, rtol=1e-12, atol=1e-14)
    assert torch.allclose((out ** 2).sum(1), torch.ones(B, dtype=out.dtype), atol=1e-9)


def test_pca_oracle_matches_sklearn():
    from sklearn.decomposition import PCA
    rng = np.random.default_rng(2)
    X = rng.standard_normal((200, 24)) @ rng.standard_normal((24, 24))
    pca = PCA(whiten=True, n_components=8).fit(X)
    v, m, var = onv.sklearn_pca_params(pca)
    y = onv.pca_project(X[:10], v, m, var).numpy()
    assert np.allclose(y, pca.transform(X[:10]), rtol=1e-9, atol=1e-10)


def test_knn_oracle_matches_reference_kdtree_call():
    rng = np.random.default_rng(3)
    ref = rng.standard_normal((600, 32)).astype(np.float32)
    qry = (ref[rng.integers(0, 600, 20)] + 0.3 * rng.standard_normal((20, 32))).astype(np.float32)
    kd_d, kd_i = orr.knn_kdtree(ref, qry, 25)                 # evaluation/top-n.py:103-106, verbatim call
    bf_d, bf_i = orr.knn_bruteforce(ref, qry, 25)
    ex_d, ex_i = orr.knn_bruteforce_exact(ref, qry, 25)
    assert np.array_equal(kd_i, bf_i) and np.array_equal(kd_i, ex_i)
    assert np.allclose(kd_d, bf_d, rtol=1e-12) and np.allclose(kd_d, ex_d, rtol=1e-12)
    sg_d, sg_i = orr.knn_sgemm_allcores(ref, qry, 25)         # the bench's all-core CPU baseline
    assert np.array_equal(kd_i, sg_i) and np.allclose(kd_d, sg_d, rtol=1e-12)


def test_recall_oracle():
    d = np.array([[30.0, 2.0, 50.0], [1.0, 9.0, 0.5], [40.0, 41.0, 3.0]])
    X, Y = orr.recall_at_n(d, rad=10.0, num=3)                 # thresholds 0, 5, 10
    assert np.allclose(X, [0, 5, 10])
    assert np.allclose(Y[0], [0, 100 / 3, 100 / 3])            # top-1: only query 1 within 5 or 10 m
    assert np.allclose(Y[1], [0, 200 / 3, 200 / 3])
    assert np.allclose(Y[2], [0, 100, 100])
    X1, Y1 = orr.recall_curve_top1(d, t=10.0, num=3)
    assert np.allclose(Y1, Y[0])


def test_wms_masks_are_float32_like_the_reference(golden):
    """mask_pos > 0 (losses.py:50) hinges on float32 overflow of tf.exp: pairs beyond ~126 m get mask_pos == 0 exactly."""
    d = torch.tensor([[0.0, 10.0, 100.0, 130.0, 500.0]])
    mp32, _ = ol.wms_masks(d.to(torch.float32), 0.8, 15.0)
    mp64, _ = ol.wms_masks(d.to(torch.float64), 0.8, 15.0)
    assert (mp32[0, :3] > 0).all() and (mp32[0, 3:] == 0).all()
    assert (mp64 > 0).all()            # which is why a float64 transcription of the masks would be a different loss
    g = golden("wms_flat_S25_D64")
    assert int(g["keptpos_exp_ms_nomine"].sum()) < 25 * 24


@pytest.mark.parametrize("n,D", [(120, 40), (60, 200)])
def test_pca_fit_oracle_matches_sklearn_full_solver(n, D):
    """the oracle's PCA fit against the reference's library call with the exact solver."""
    from sklearn.decomposition import PCA
    X = synth.pca_features(n, D, rank=12, seed=5)
    pca = PCA(whiten=True, n_components=10, svd_solver="full").fit(X)
    v, m, var = onv.pca_fit(X, 10)
    assert np.allclose(m, pca.mean_, rtol=1e-12, atol=1e-12)
    assert np.allclose(var, pca.explained_variance_, rtol=1e-10)
    assert np.allclose(v, pca.components_, rtol=1e-8, atol=1e-9)
    assert np.allclose(onv.pca_project(X[:7], v, m, var).numpy(), pca.transform(X[:7]), rtol=1e-8, atol=1e-9)
