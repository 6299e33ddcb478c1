"""GPU parity: exact kNN (both device paths), shard merge, geo bookkeeping, recall -- indices bit-exact."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from oracle import retrieval as orr
from soft_contrastive_learning_b200 import synth

pytestmark = pytest.mark.gpu


def check_exact(d, i, rd, ri):
    assert np.array_equal(i, ri), f"{(i != ri).sum()} index mismatches"
    assert np.allclose(d, rd, rtol=1e-12, atol=0)
    assert (np.diff(d, axis=1) >= 0).all()


@pytest.mark.parametrize("R,Q,D,k", [(600, 20, 32, 25), (1000, 1, 4096, 1000), (5000, 64, 256, 5), (40, 3, 8, 25)])
def test_exact_scan_path(cuda_lib, R, Q, D, k):
    from soft_contrastive_learning_b200 import retrieval
    db, qry, *_ = synth.retrieval_problem(R=R, Q=Q, D=D, seed=1)
    tree = retrieval.KDTree(db)
    d, i = tree.query(qry, k=k, force_path=1)
    rd, ri = orr.knn_bruteforce_exact(db, qry, k)
    kk = min(k, R)
    check_exact(d[:, :kk], i[:, :kk], rd, ri)
    if k > R:
        assert np.isinf(d[:, R:]).all() and (i[:, R:] == -1).all()
    assert tree.stats()["path"] == 1


def test_exact_scan_matches_reference_kdtree_call(cuda_lib):
    from soft_contrastive_learning_b200 import retrieval
    db, qry, *_ = synth.retrieval_problem(R=3000, Q=32, D=64, seed=2)
    kd_d, kd_i = orr.knn_kdtree(db, qry, 25)                 # evaluation/top-n.py:103-106
    d, i = retrieval.KDTree(db).query(qry, k=25, return_distance=True, sort_results=True)
    check_exact(d, i, kd_d, kd_i)


def test_ties_are_ordered_by_index(cuda_lib):
    from soft_contrastive_learning_b200 import retrieval
    rng = np.random.default_rng(3)
    base = rng.standard_normal((50, 16)).astype(np.float32)
    db = np.concatenate([base, base, base], 0)               # every row appears three times
    qry = base[:7] + 0.0
    for path in (1,):
        d, i = retrieval.KDTree(db).query(qry, k=6, force_path=path)
        rd, ri = orr.knn_bruteforce_exact(db, qry, 6)
        check_exact(d, i, rd, ri)
        assert (i[:, :3] == np.arange(7)[:, None] + np.array([0, 50, 100])[None]).all()


@pytest.fixture(params=["1", "2", "3"], ids=["cta_group1", "cta_pair", "cta_pair_wide"])
def tc_variant(request, monkeypatch):
    """Both tensor-pass kernels: single-CTA 128x256 tiles and cta_group::2 CTA pairs (256x256)."""
    monkeypatch.setenv("SCL_KNN_TC_VARIANT", request.param)
    return request.param


def test_tensor_pass_raw_scores_match_fp16_gemm(cuda_lib, tc_variant):
    """The tcgen05 GEMM itself: scores |r|^2 - 2 q.r from the fp16 pass vs the same fp16-rounded operands in float64."""
    from soft_contrastive_learning_b200 import retrieval
    R, Q, D = 4096 + 300, 150, 192
    db, qry, *_ = synth.retrieval_problem(R=R, Q=Q, D=D, seed=4)
    tree = retrieval.KDTree(db)
    dbg = torch.full((Q, R), float("nan"), dtype=torch.float32, device="cuda")
    os.environ["SCL_KNN_DEBUG_SCORES"] = hex(dbg.data_ptr())
    try:
        tree.query_device(torch.tensor(qry, device="cuda"), k=25, force_path=2)
    finally:
        del os.environ["SCL_KNN_DEBUG_SCORES"]
    torch.cuda.synchronize()
    got = dbg.cpu().numpy().astype(np.float64)
    assert not np.isnan(got).any(), "some tiles were never written"

    e_db = 13 - int(np.floor(np.log2(np.abs(db).max())))
    dbh = (db * 2.0 ** e_db).astype(np.float16).astype(np.float64) * 2.0 ** -e_db
    e_q = 13 - np.floor(np.log2(np.abs(qry).max(axis=1)))
    qh = (qry * (2.0 ** e_q)[:, None]).astype(np.float16).astype(np.float64) * (2.0 ** -e_q)[:, None]
    rn = (db.astype(np.float64) ** 2).sum(1).astype(np.float32).astype(np.float64)
    want = rn[None, :] - 2.0 * (qh @ dbh.T)
    err = np.abs(got - want).max()
    assert err < 2e-3 * np.abs(want).max() * 1e-2, err      # fp32 accumulation / epilogue rounding only


@pytest.mark.parametrize("R,Q,D,k", [(20000, 300, 256, 25), (9000, 129, 4096, 25), (50000, 64, 128, 5),
                                     (4097, 257, 64, 32)])
def test_tensor_pass_is_exact(cuda_lib, tc_variant, R, Q, D, k):
    from soft_contrastive_learning_b200 import retrieval
    db, qry, *_ = synth.retrieval_problem(R=R, Q=Q, D=D, seed=5)
    tree = retrieval.KDTree(db)
    d, i = tree.query(qry, k=k, force_path=2)
    st = tree.stats()
    assert st["path"] == 2 and st["n_certified"] + st["n_fallback"] == Q
    rd, ri = orr.knn_bruteforce(db, qry, k)
    check_exact(d, i, rd, ri)
    # Gaussian descriptors leave a wide gap between rank k and rank 64: (almost) everything certifies
    assert st["n_fallback"] <= Q // 10, st


def test_fallback_path_is_exact(cuda_lib):
    from soft_contrastive_learning_b200 import retrieval
    db, qry, *_ = synth.retrieval_problem(R=6000, Q=40, D=128, seed=6)
    tree = retrieval.KDTree(db)
    d, i = tree.query(qry, k=25, force_path=3)               # every query forced through the exact-scan fallback
    assert tree.stats()["n_fallback"] == 40
    rd, ri = orr.knn_bruteforce(db, qry, 25)
    check_exact(d, i, rd, ri)


def test_clustered_descriptors_trigger_the_certificate(cuda_lib):
    """Adversarial for fp16: 200 near-duplicates of each query direction differ by less than the rounding bound,
    so the certificate must refuse and the exact path must still return the right order."""
    from soft_contrastive_learning_b200 import retrieval
    rng = np.random.default_rng(7)
    D = 128
    centers = rng.standard_normal((8, D)).astype(np.float32)
    near = (centers[:, None, :] + 1e-4 * rng.standard_normal((8, 200, D))).reshape(-1, D).astype(np.float32)
    far = rng.standard_normal((4000, D)).astype(np.float32)
    db = np.concatenate([far, near], 0)
    qry = centers
    tree = retrieval.KDTree(db)
    d, i = tree.query(qry, k=25, force_path=2)
    rd, ri = orr.knn_bruteforce_exact(db, qry, 25)
    check_exact(d, i, rd, ri)
    assert tree.stats()["n_fallback"] >= 1


def test_topk_merge_and_index_offsets(cuda_lib):
    from soft_contrastive_learning_b200 import retrieval
    db, qry, *_ = synth.retrieval_problem(R=3001, Q=17, D=64, seed=8)
    G, k = 4, 25
    ds, is_ = [], []
    for r in range(G):
        lo, hi = retrieval.shard_bounds(3001, G, r)
        d, i = retrieval.KDTree(db[lo:hi], index_offset=lo).query_device(torch.tensor(qry, device="cuda"), k)
        ds.append(d)
        is_.append(i)
    d, i = retrieval.topk_merge(torch.stack(ds), torch.stack(is_))
    rd, ri = orr.knn_bruteforce_exact(db, qry, k)
    check_exact(d.cpu().numpy(), i.cpu().numpy(), rd, ri)


def test_geo_and_recall_and_top_n_payload(cuda_lib):
    from soft_contrastive_learning_b200 import retrieval
    db, qry, ref_xy, query_xy, _ = synth.retrieval_problem(R=2500, Q=60, D=64, seed=9, extent=300.0)
    for l in (0.0, 5.0):
        got = retrieval.top_n(db, qry, ref_xy, query_xy, N=25, l=l)
        ref = orr.top_n(db, qry, ref_xy, query_xy, N=25, l=l)
        assert got[5] == ref[5]                                              # ref_idx
        assert np.array_equal(np.asarray(got[0]), np.asarray(ref[0]))        # top_i (original indices)
        assert np.allclose(np.asarray(got[1]), np.asarray(ref[1]), rtol=0, atol=1e-6)   # top_g_dists (sklearn expands)
        assert np.allclose(got[2], ref[2], rtol=1e-12)                       # top_f_dists
        assert np.array_equal(np.asarray(got[3]), np.asarray(ref[3]))        # gt_i
        assert np.allclose(got[4], ref[4], atol=1e-6)
        X, Y = retrieval.recall_at_n(np.asarray(got[1]), rad=25.0, num=25)
        Xo, Yo = orr.recall_at_n(np.asarray(got[1]), rad=25.0, num=25)
        assert np.array_equal(X, Xo) and np.array_equal(Y, Yo)
        Xo1, Yo1 = orr.recall_curve_top1(got[1], t=25.0, num=50)
        assert np.array_equal(retrieval.recall_curves(np.asarray(got[1]), Xo1)[0], Yo1)


def test_full_size_properties_1M_x_4096(cuda_lib):
    """BASELINE config 4 size (1M x 4096 fp32 database) through size-independent properties:
    queries that ARE database rows return themselves at distance 0; lists are sorted; the tensor pass and the
    exact scan agree bit-for-bit on a query subset."""
    from soft_contrastive_learning_b200 import retrieval
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2 ** 30:
        pytest.skip("needs ~30 GB of HBM")
    R, D, Q = 1_000_000, 4096, 512
    g = torch.Generator(device="cuda").manual_seed(42)
    db = torch.randn((R, D), generator=g, device="cuda", dtype=torch.float32)
    src = torch.randint(0, R, (Q,), generator=g, device="cuda")
    qry = db[src] + 0.5 * torch.randn((Q, D), generator=g, device="cuda")
    qry[:64] = db[src[:64]]
    tree = retrieval.KDTree(db)
    d, i = tree.query_device(qry, k=25)
    st = tree.stats()
    assert st["path"] == 2
    assert (i[:, 0] == src).all()                                  # planted neighbour is rank 1
    assert (d[:64, 0] == 0).all()
    assert (d[:, 1:] >= d[:, :-1]).all()
    d2, i2 = tree.query_device(qry[:12], k=25, force_path=1)       # exact scan on a subset
    assert torch.equal(i[:12], i2) and torch.equal(d[:12], d2)
    assert st["n_fallback"] <= Q // 20, st


# ---------------------------------------------------------------------------------------------
# the callers either side of the kernels (files in / pickles out, mining, localization)
# ---------------------------------------------------------------------------------------------
def _write_eval_files(tmp_path, R=1500, Q=40, D=96, seed=4):
    from soft_contrastive_learning_b200 import formats
    db, qry, ref_xy, query_xy, _ = synth.retrieval_problem(R=R, Q=Q, D=D, seed=seed, extent=200.0)
    rng = np.random.default_rng(seed)
    # drive along a path so that greedy subsampling by distance (top-n.py:91-94) actually drops references
    order = np.argsort(ref_xy[:, 0])
    db, ref_xy = db[order], ref_xy[order]
    pca_f = db[rng.choice(R, 400, replace=False)] + 0.01 * rng.standard_normal((400, D)).astype(np.float32)
    paths = {k: str(tmp_path / (k + ".x")) for k in ("pca", "ref", "query")}
    formats.save_features(pca_f, paths["pca"] + ".pickle")
    formats.save_features(db, paths["ref"] + ".pickle")
    formats.save_features(qry, paths["query"] + ".v1.pickle")
    for name, xy in (("ref", ref_xy), ("query", query_xy)):
        formats.save_csv({"t": list(range(len(xy))), "easting": [repr(float(v)) for v in xy[:, 0]],
                          "northing": [repr(float(v)) for v in xy[:, 1]]}, paths[name] + ".csv")
    return paths, pca_f, db, qry, ref_xy, query_xy


@pytest.mark.parametrize("pca_solver", ["sklearn", "gpu"])
def test_get_top_n_files_in_pickles_out(cuda_lib, tmp_path, pca_solver):
    from sklearn.decomposition import PCA
    from soft_contrastive_learning_b200 import evaluation, formats, netvlad
    paths, pca_f, db, qry, ref_xy, query_xy = _write_eval_files(tmp_path)
    out_root = str(tmp_path / "top_n")
    L, Dm = (0.0, 3.0), (32, 64)
    written = evaluation.get_top_n(paths["pca"] + ".pickle", paths["query"] + ".v1.pickle", paths["ref"] + ".pickle",
                                   paths["query"] + ".csv", paths["ref"] + ".csv", out_root, N=25, L=L, D=Dm,
                                   log=lambda *_: None, pca_solver=pca_solver)
    assert len(written) == 4
    if pca_solver == "gpu":
        v_all, m_all, var_all = netvlad.pca_fit(pca_f, max(Dm))               # 8f row 4: one exact fit for the sweep
    for d in Dm:
        if pca_solver == "gpu":
            v, m, var = v_all[:d], m_all, var_all[:d]
            pca = PCA(whiten=True, n_components=d, svd_solver="full").fit(pca_f.astype(np.float64))
            assert np.allclose(var, pca.explained_variance_, rtol=1e-4)
        else:
            pca = PCA(whiten=True, n_components=d).fit(pca_f)                 # the reference's host pipeline, top-n.py:74-77
            v, m, var = netvlad.pca_from_sklearn(pca)
        pr, pq = netvlad.pca_project(db, v, m, var), netvlad.pca_project(qry, v, m, var)
        # P1 eval twin: the device projection is sklearn's transform to fp32 rounding; the neighbour lists below are then
        # compared on identical projected features, so that they must agree exactly (no near-tie reordering)
        if pca_solver == "sklearn":
            assert np.allclose(pr, pca.transform(db), rtol=0, atol=2e-5 * np.abs(pr).max())
        for l in L:
            f = os.path.join(out_root, "l{}_dim{}".format(l, d), "queryxv1.pickle")   # name rule of top-n.py:84
            assert f in written
            got = formats.load_pickle(f)
            ref = orr.top_n(pr, pq, ref_xy, query_xy, N=25, l=l)
            assert len(got) == 6 and got[5] == ref[5]
            assert np.array_equal(np.asarray(got[0]), np.asarray(ref[0]))     # same neighbours, original indices
            assert np.allclose(got[2], ref[2], rtol=1e-12)
            assert np.allclose(np.asarray(got[1]), np.asarray(ref[1]), atol=1e-6)
            assert np.array_equal(np.asarray(got[3]), np.asarray(ref[3])) and np.allclose(got[4], ref[4], atol=1e-6)
    # second call: everything exists -> nothing recomputed (top-n.py:41-57)
    assert evaluation.get_top_n(paths["pca"] + ".pickle", paths["query"] + ".v1.pickle", paths["ref"] + ".pickle",
                                paths["query"] + ".csv", paths["ref"] + ".csv", out_root, N=25, L=L, D=Dm,
                                log=lambda *_: None, pca_solver=pca_solver) == []


def test_mining_cache_full_sort(cuda_lib):
    from soft_contrastive_learning_b200 import evaluation
    rng = np.random.default_rng(12)
    feats = rng.standard_normal((1000, 256)).astype(np.float32)             # MINING_CACHE_SIZE images
    idx = rng.permutation(50000)[:1000]
    cache = evaluation.FeatureCache(feats, idx)
    for index in (int(idx[0]), int(idx[517])):
        got = cache.sorted_neighbours(index)
        ref = orr.mining_sorted_neighbours(feats, idx, index, k=1000)
        assert got[0] == index and [int(g) for g in got] == [int(r) for r in ref]
    assert cache.sorted_neighbours(-5) is None                               # not cached: train.py:447


def test_in_training_localization(cuda_lib):
    from sklearn.neighbors import KDTree as SkKDTree
    from soft_contrastive_learning_b200 import evaluation
    db, qry, ref_xy, query_xy, _ = synth.retrieval_problem(R=3000, Q=100, D=128, seed=6, extent=400.0)
    ld, li, gd, gi = evaluation.evaluate_localization(db, qry, ref_xy, query_xy, k=5)
    rd, ri = SkKDTree(db).query(qry, k=5)                                    # train.py:1181-1182
    od, oi = SkKDTree(ref_xy).query(query_xy, k=1)                           # train.py:1184-1185
    assert np.array_equal(li, ri) and np.allclose(ld, rd, rtol=1e-12)
    assert np.array_equal(gi, oi) and np.allclose(gd, od, rtol=1e-12)
    got = evaluation.localization_summary(li, gd, query_xy, ref_xy)
    ref = orr.localization_summary(ri, od, query_xy, ref_xy)
    for rad in (50, 25, 10):
        assert np.array_equal(got["curves"][rad]["Y"], ref["curves"][rad]["Y"])
        assert np.array_equal(got["curves"][rad]["optimum"], ref["curves"][rad]["optimum"])
    for tag, v in ref["scalars"].items():
        assert abs(got["scalars"][tag] - v) <= 1e-9 * max(1.0, abs(v)), tag
