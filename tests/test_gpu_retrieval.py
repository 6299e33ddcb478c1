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
    d, i = retrieval.KDTree(db).query(qry, k=6, force_path=1)
    rd, ri = orr.knn_bruteforce_exact(db, qry, 6)
    check_exact(d, i, rd, ri)
    assert (i[:, :3] == np.arange(7)[:, None] + np.array([0, 50, 100])[None]).all()


@pytest.mark.parametrize("force_path", [2, 4], ids=["tensor_pass", "tensor_pass_then_stage2"])
@pytest.mark.parametrize("copies,k", [(3, 6), (3, 25), (40, 25), (100, 25)])
def test_exact_ties_on_the_tensor_path(cuda_lib, force_path, copies, k):
    """Exact duplicate rows through the tcgen05 candidate pass: equal fp16 scores, the `s_k + 2 eps` cut, the (distance,
    index) order of the rescored candidates and -- with 40 / 100 copies -- more ties than the k' = 64 candidate list
    holds, which the certificate must refuse and the second tensor stage (or the exact scan) must order by index."""
    from soft_contrastive_learning_b200 import retrieval
    rng = np.random.default_rng(13)
    n_base = 6000 // copies
    base = rng.standard_normal((n_base, 64)).astype(np.float32)
    db = np.concatenate([base] * copies, 0)                  # row j appears at j, j + n_base, j + 2 n_base, ...
    nq = min(96, n_base)
    qry = np.concatenate([base[:nq], base[:32] + 0.25 * rng.standard_normal((32, 64)).astype(np.float32)], 0)
    tree = retrieval.KDTree(db)
    d, i = tree.query(qry, k=k, force_path=force_path)
    assert tree.stats()["path"] == 2
    rd, ri = orr.knn_bruteforce_exact(db, qry, k)
    check_exact(d, i, rd, ri)
    m = min(copies, k)
    assert (i[:nq, :m] == np.arange(nq)[:, None] + n_base * np.arange(m)[None]).all()
    assert (d[:nq, :m] == 0).all()


def test_exact_ties_across_shards_and_the_merge(cuda_lib):
    """Duplicates that live on DIFFERENT shards: per-shard tensor passes, global index offsets, the merge kernel's
    (distance, index) order -- through the packed one-message layout the NCCL path uses."""
    from soft_contrastive_learning_b200 import retrieval
    rng = np.random.default_rng(14)
    base = rng.standard_normal((1500, 64)).astype(np.float32)
    db = np.concatenate([base] * 4, 0)                       # 6000 rows, copy c of row j at c * 1500 + j
    qry = base[:50]
    G, k, Q = 3, 25, 50                                      # 2000 rows per shard: the copies straddle shard borders
    packed = torch.empty((G, 2, Q, k), dtype=torch.int64, device="cuda")
    for r in range(G):
        lo, hi = retrieval.shard_bounds(6000, G, r)
        retrieval.KDTree(db[lo:hi], index_offset=lo).query_device(torch.tensor(qry, device="cuda"), k, force_path=2,
                                                                  out=(packed[r, 0].view(torch.float64), packed[r, 1]))
    d, i = retrieval.topk_merge_packed(packed, G, Q, k)
    rd, ri = orr.knn_bruteforce_exact(db, qry, k)
    check_exact(d.cpu().numpy(), i.cpu().numpy(), rd, ri)
    assert (i[:, :4].cpu().numpy() == np.arange(50)[:, None] + 1500 * np.arange(4)[None]).all()


def _two_phase(db, qry, G, k, bounds=None, pipelined=False):
    """G shards of `db` on ONE GPU through the two-phase protocol of retrieval.ShardedKDTree with the all-gathers replaced
    by writes into one [G, ...] buffer: scl_knn_query_begin -> scl_knn_bound_reduce -> scl_knn_query_end -> merge, or
    (pipelined) scl_knn_query_launch on the current stream and, per query group on a SECOND stream,
    scl_knn_query_begin_group (stream wait on the tensor kernel's signal) -> bound reduce -> scl_knn_query_end_group."""
    from soft_contrastive_learning_b200 import retrieval
    R, Q = db.shape[0], qry.shape[0]
    q = torch.tensor(qry, device="cuda")
    trees = []
    for r in range(G):
        lo, hi = bounds[r] if bounds else retrieval.shard_bounds(R, G, r)
        trees.append(retrieval.KDTree(db[lo:hi], index_offset=lo))
    d = torch.empty((Q, k), dtype=torch.float64, device="cuda")
    i = torch.empty((Q, k), dtype=torch.int64, device="cuda")
    real, n_groups = 0, 1
    if not pipelined:
        ub_all = torch.empty((G, Q, k), dtype=torch.float32, device="cuda")
        states = [trees[r].query_begin(q, k, ub_all[r]) for r in range(G)]
        bound = retrieval.bound_reduce(ub_all)
        assert torch.equal(bound, ub_all.permute(1, 0, 2).reshape(Q, G * k).sort(1).values[:, k - 1])
        packed = torch.empty((G, 2, Q, k), dtype=torch.int64, device="cuda")
        for r in range(G):
            out = (packed[r, 0].view(torch.float64), packed[r, 1])
            if states[r] is None:
                trees[r].query_device(q, k, 0, out=out)
            else:
                trees[r].query_end(states[r], q, k, bound, out)
        retrieval.topk_merge_packed(packed, G, Q, k, out=(d, i))
        real = (packed[:, 1] >= 0).sum().item()
    else:
        n_groups, gq = retrieval.KDTree.query_groups(db.shape[1], Q)
        main, side = torch.cuda.current_stream(), torch.cuda.Stream()
        states = [trees[r].query_launch(q, k) for r in range(G)]          # G persistent tensor launches queued on `main`
        if any(st is None for st in states):
            side.wait_stream(main)
        with torch.cuda.stream(side):
            for g in range(n_groups):
                q0, nq = g * gq, min(gq, Q - g * gq)
                ub_all = torch.empty((G, nq, k), dtype=torch.float32, device="cuda")
                for r in range(G):
                    if states[r] is None:
                        ub_all[r].fill_(float("inf"))
                    else:
                        trees[r].query_begin_group(states[r], q, k, g, ub_all[r])
                bound = retrieval.bound_reduce(ub_all)
                packed = torch.empty((G, 2, nq, k), dtype=torch.int64, device="cuda")
                for r in range(G):
                    out = (packed[r, 0].view(torch.float64), packed[r, 1])
                    if states[r] is None:
                        trees[r].query_device(q[q0:q0 + nq], k, 0, out=out)
                    else:
                        trees[r].query_end_group(states[r], q, k, g, bound, out)
                retrieval.topk_merge_packed(packed, G, nq, k, out=(d[q0:q0 + nq], i[q0:q0 + nq]))
                real += (packed[:, 1] >= 0).sum().item()
        main.wait_stream(side)
    torch.cuda.synchronize()
    return d.cpu().numpy(), i.cpu().numpy(), real, [t.stats() for t in trees], n_groups


@pytest.mark.parametrize("R,Q,D,k,G", [(40000, 300, 256, 25, 4), (24000, 129, 4096, 25, 3), (64000, 64, 128, 5, 8),
                                       (9000, 257, 64, 32, 2)])
@pytest.mark.parametrize("pipelined", [False, True], ids=["begin_end", "group_pipeline"])
def test_two_phase_sharded_query_is_exact(cuda_lib, tune, R, Q, D, k, G, pipelined):
    """Between the two phases the shards agree on a bound on the k-th global distance; each rescoring only what can
    still reach the global top-k.  Result = float64 brute force over the whole database, and the shards together return
    far fewer than G*k rows per query."""
    db, qry, *_ = synth.retrieval_problem(R=R, Q=Q, D=D, seed=15)
    if pipelined:
        tune("SCL_KNN_GROUP_M", 1)                               # one work unit (256 queries) per group: several groups
    d, i, real, stats, n_groups = _two_phase(db, qry, G, k, pipelined=pipelined)
    rd, ri = orr.knn_bruteforce(db, qry, k)
    check_exact(d, i, rd, ri)
    assert Q * k <= real <= 2 * Q * k, (real, Q * k, G)          # plain protocol: G * Q * k
    assert n_groups == (-(-Q // 256) if pipelined else 1)
    for st in stats:
        assert st["path"] == 2 and st["n_certified"] + st["n_fallback"] == Q, st


@pytest.mark.parametrize("pipelined", [False, True], ids=["begin_end", "group_pipeline"])
def test_two_phase_ties_duplicates_and_refusals(cuda_lib, tune, pipelined):
    """Adversarial for the bound: exact duplicates that straddle shard borders (equal scores at the cut; the row left out
    must be STRICTLY farther), near-duplicate clusters living on ONE shard (that shard has more than 64 rows inside the
    bound: it falls back to its exact local top-k through the second tensor stage / scan while the others trim), a shard
    too small for the tensor pass (contributes +inf, answers with the plain query) and a ragged last shard."""
    from soft_contrastive_learning_b200 import retrieval
    rng = np.random.default_rng(16)
    D, k = 128, 25
    base = rng.standard_normal((3000, D)).astype(np.float32)
    centers = rng.standard_normal((6, D)).astype(np.float32)
    near = (centers[:, None, :] + 1e-4 * rng.standard_normal((6, 150, D))).reshape(-1, D).astype(np.float32)
    db = np.concatenate([base, base, near, base[:1700], rng.standard_normal((500, D)).astype(np.float32)], 0)
    qry = np.concatenate([base[:40], centers, base[100:120] + 0.2 * rng.standard_normal((20, D)).astype(np.float32)], 0)
    R = db.shape[0]                                                  # 9100 rows
    bounds = [(0, 2500), (2500, 5000), (5000, 8600), (8600, 9100)]   # last shard: 500 rows < 1024 -> unsupported
    if pipelined:
        tune("SCL_KNN_GROUP_M", 1)
    d, i, real, stats, _ = _two_phase(db, qry, 4, k, bounds, pipelined=pipelined)
    rd, ri = orr.knn_bruteforce_exact(db, qry, k)
    check_exact(d, i, rd, ri)
    assert (i[:40, :3] == np.arange(40)[:, None] + np.array([0, 3000, 6900])[None]).all() and (d[:40, :3] == 0).all()
    assert stats[3]["path"] == 1                                     # the small shard took the plain exact scan
    assert stats[2]["n_fallback"] >= 6, stats[2]                     # the cluster shard refused the cluster queries
    # the same database through ONE index gives the same bits
    tune("SCL_KNN_GROUP_M", None)
    d1, i1 = retrieval.KDTree(db).query(qry, k=k, force_path=2)
    assert np.array_equal(i1, i) and np.array_equal(d1, d)


@pytest.mark.parametrize("pipelined", [False, True], ids=["begin_end", "group_pipeline"])
def test_two_phase_large_shards_match_single_index(cuda_lib, pipelined):
    """Config-4-like sizes (D = 4096; 6000 queries = two query groups of the tensor launch) on 2 shards of one GPU: the
    second stream merges / rescoring group 0 while the tensor kernel works on group 1."""
    from soft_contrastive_learning_b200 import retrieval
    db, qry, *_ = synth.retrieval_problem(R=60000, Q=6000, D=4096, seed=17)
    d, i, real, stats, n_groups = _two_phase(db, qry, 2, 25, pipelined=pipelined)
    assert n_groups == (2 if pipelined else 1) and stats[0]["chunks"] == n_groups
    d1, i1 = retrieval.KDTree(db).query(qry, k=25)
    assert np.array_equal(i1, i) and np.array_equal(d1, d)
    assert real <= 0.85 * 2 * 6000 * 25                       # plain protocol: 2 * Q * k; here ~25 + the rows within 2 eps


@pytest.fixture(params=["1", "2", "3"], ids=["cta_group1", "cta_pair", "cta_pair_wide"])
def tc_variant(request, tune):
    """The tensor-pass kernels: single-CTA 128x256 tiles and cta_group::2 CTA pairs (256x256, 256x512)."""
    tune("SCL_KNN_TC_VARIANT", int(request.param))
    return request.param


def test_tensor_pass_raw_scores_match_fp16_gemm(cuda_lib, tc_variant):
    """The tcgen05 GEMM itself: scores |r|^2 - 2 q.r from the fp16 pass vs the same fp16-rounded operands in float64."""
    from soft_contrastive_learning_b200 import retrieval
    R, Q, D = 4096 + 300, 150, 192
    db, qry, *_ = synth.retrieval_problem(R=R, Q=Q, D=D, seed=4)
    tree = retrieval.KDTree(db)
    dbg = torch.full((Q, R), float("nan"), dtype=torch.float32, device="cuda")
    cuda_lib.scl_knn_set_debug_scores(C.c_void_p(dbg.data_ptr()), dbg.numel())    # explicit, size-checked test hook
    try:
        tree.query_device(torch.tensor(qry, device="cuda"), k=25, force_path=2)
    finally:
        cuda_lib.scl_knn_set_debug_scores(None, 0)
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


@pytest.mark.parametrize("force_path,key", [(3, "n_scan"), (4, "n_stage2")], ids=["exact_scan", "stage2"])
def test_fallback_paths_are_exact(cuda_lib, force_path, key):
    from soft_contrastive_learning_b200 import retrieval
    db, qry, *_ = synth.retrieval_problem(R=6000, Q=40, D=128, seed=6)
    tree = retrieval.KDTree(db)
    d, i = tree.query(qry, k=25, force_path=force_path)      # every query forced through the exact scan / the second stage
    st = tree.stats()
    assert st["n_fallback"] == 40 and st[key] == 40, st
    rd, ri = orr.knn_bruteforce(db, qry, 25)
    check_exact(d, i, rd, ri)


@pytest.mark.parametrize("dups,expect", [(200, "n_stage2"), (3000, "n_scan")])
def test_clustered_descriptors_trigger_the_certificate(cuda_lib, dups, expect):
    """Adversarial for fp16: `dups` near-duplicates of each query direction differ by less than the rounding bound, so the
    certificate must refuse.  200 of them fit the second tensor stage's list (2048 rows inside the bound); 3000 overflow it
    and go to the float64 scan.  Either way the order is the exact one."""
    from soft_contrastive_learning_b200 import retrieval
    rng = np.random.default_rng(7)
    D = 128
    centers = rng.standard_normal((8, D)).astype(np.float32)
    near = (centers[:, None, :] + 1e-4 * rng.standard_normal((8, dups, D))).reshape(-1, D).astype(np.float32)
    far = rng.standard_normal((4000, D)).astype(np.float32)
    db = np.concatenate([far, near], 0)
    qry = centers
    tree = retrieval.KDTree(db)
    d, i = tree.query(qry, k=25, force_path=2)
    rd, ri = orr.knn_bruteforce_exact(db, qry, 25)
    check_exact(d, i, rd, ri)
    st = tree.stats()
    assert st["n_fallback"] == 8 and st[expect] == 8, st


def test_stage2_resolves_refused_queries_of_a_trajectory(cuda_lib, tune):
    """Clustered descriptors as a real traversal produces them (synth.trajectory_problem: AR(1) frames with stops where
    the vehicle stands still and hundreds of frames are near-identical).  Queries that fall on a stop are refused by the
    first pass and must be resolved by the second tensor stage, not by the float64 scan; with stage 2 switched off the
    scan gives the same answer.  Checked against the reference's own call, sklearn KDTree.query."""
    from soft_contrastive_learning_b200 import retrieval
    db, qry, info = synth.trajectory_problem(R=12000, Q=256, D=128, seed=3, stop_frac=0.25, stop_len=(150, 400))
    tree = retrieval.KDTree(db)
    d, i = tree.query(qry, k=25, force_path=2)
    st = tree.stats()
    kd_d, kd_i = orr.knn_kdtree(db, qry, 25)                 # evaluation/top-n.py:103-106
    check_exact(d, i, kd_d, kd_i)
    assert st["n_fallback"] >= 16 and st["n_scan"] == 0 and st["n_stage2"] == st["n_fallback"], st
    tune("SCL_KNN_STAGE2", 0)
    d0, i0 = tree.query(qry, k=25, force_path=2)
    assert tree.stats()["n_scan"] == st["n_fallback"]
    assert np.array_equal(i0, i) and np.array_equal(d0, d)


@pytest.mark.parametrize("chunk", [256, 512])
def test_chunk_pipeline_matches_single_chunk(cuda_lib, tune, chunk):
    """scl_knn_query splits the queries into chunks and runs the merge / rescore / certificate of one chunk on a helper
    stream under the tensor pass of the next: any chunking must return bit-identical results (ragged last chunk,
    refused queries in several chunks)."""
    from soft_contrastive_learning_b200 import retrieval
    db, qry, info = synth.trajectory_problem(R=9000, Q=1100, D=64, seed=5, stop_frac=0.1, stop_len=(100, 200))
    tree = retrieval.KDTree(db)
    tune("SCL_KNN_CHUNK_Q", 0)
    d1, i1 = tree.query(qry, k=25, force_path=2)
    assert tree.stats()["chunks"] == 1
    tune("SCL_KNN_CHUNK_Q", chunk)
    d2, i2 = tree.query(qry, k=25, force_path=2)
    assert tree.stats()["chunks"] == -(-1100 // chunk)
    assert np.array_equal(i1, i2) and np.array_equal(d1, d2)
    rd, ri = orr.knn_bruteforce_exact(db, qry, 25)
    check_exact(d2, i2, rd, ri)


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


def _fp64_topk_chunked(db, qry, k, chunk=32768):
    """Independent float64 reference on the device, nothing of the product in it: d^2 = |q|^2 + |r|^2 - 2 q.r with a
    float64 GEMM per row chunk, torch.topk per chunk, final order by (d^2, index) on the host."""
    q64 = qry.double()
    qn = (q64 * q64).sum(1, keepdim=True)
    cand_d, cand_i = [], []
    for r0 in range(0, db.shape[0], chunk):
        r64 = db[r0:r0 + chunk].double()
        d2 = qn + (r64 * r64).sum(1)[None, :] - 2.0 * (q64 @ r64.t())
        v, ix = torch.topk(d2, min(k + 8, d2.shape[1]), dim=1, largest=False)
        cand_d.append(v.cpu().numpy())
        cand_i.append((ix + r0).cpu().numpy())
    cd, ci = np.concatenate(cand_d, 1), np.concatenate(cand_i, 1)
    out_d, out_i = np.empty((qry.shape[0], k)), np.empty((qry.shape[0], k), dtype=np.int64)
    for q in range(qry.shape[0]):
        order = np.lexsort((ci[q], cd[q]))[:k]
        out_d[q], out_i[q] = cd[q][order], ci[q][order]
    return out_d, out_i


def test_full_size_properties_1M_x_4096(cuda_lib, measured):
    """BASELINE config 4 size (1M x 4096 fp32 database): 512 queries against an INDEPENDENT chunked float64 computation
    (torch float64 GEMM + topk on the device, no kernel of this repo), plus size-independent properties: queries that ARE
    database rows return themselves at distance 0, lists are sorted, the pipelined chunks agree with one chunk."""
    from soft_contrastive_learning_b200 import _lib, retrieval
    free, _ = torch.cuda.mem_get_info()
    if free < 60 * 2 ** 30:
        pytest.skip("needs ~30 GB of HBM")
    R, D, Q = 1_000_000, 4096, 512
    g = torch.Generator(device="cuda").manual_seed(42)
    db = torch.empty((R, D), device="cuda", dtype=torch.float32)
    for r0 in range(0, R, 65536):
        db[r0:r0 + 65536] = torch.randn((min(65536, R - r0), D), generator=g, device="cuda")
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
    ref_d2, ref_i = _fp64_topk_chunked(db, qry, 25)
    assert np.array_equal(i.cpu().numpy(), ref_i), f"{(i.cpu().numpy() != ref_i).sum()} index mismatches vs float64 GEMM"
    got_d2 = d.cpu().numpy() ** 2
    # the GEMM form cancels |q|^2 + |r|^2 ~ 8200 down to d^2: absolute error ~ 8200 * 2^-52 * sqrt(4096)
    err = np.abs(got_d2[:, 1:] - ref_d2[:, 1:]).max()
    measured("retrieval_1Mx4096_vs_fp64_gemm", max_abs_d2_err=err, n_queries=Q, n_fallback=st["n_fallback"])
    assert err < 1e-8
    assert st["n_fallback"] <= Q // 20, st
    with _lib.tuning(SCL_KNN_CHUNK_Q=128):                         # rounded up to the 256-query tile unit: 2 pipelined chunks
        d4, i4 = tree.query_device(qry, k=25)
    assert tree.stats()["chunks"] == 2 and torch.equal(i4, i) and torch.equal(d4, d)


def _nccl_worker(rank, world, port, out):
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    try:
        from soft_contrastive_learning_b200 import retrieval
        R, Q, D, k = 40000, 300, 128, 25
        db, qry, info = synth.trajectory_problem(R=R, Q=Q, D=D, seed=9, stop_frac=0.05, stop_len=(100, 200))
        db = np.concatenate([db, db[:2000]], 0)                     # exact duplicates that land on the OTHER shard
        lo, hi = retrieval.shard_bounds(db.shape[0], world, rank)
        tree = retrieval.ShardedKDTree(torch.tensor(db[lo:hi], device="cuda"), index_offset=lo)
        d, i = tree.query_device(torch.tensor(qry, device="cuda"), k)
        d2, i2 = tree.query_from_host(torch.tensor(qry).pin_memory(), k)
        assert torch.equal(i, i2) and torch.equal(d, d2)
        # the three protocol forms give the same bits: two-phase (default), two-phase per query group on a second
        # stream (groups of 256 queries here), single phase
        from soft_contrastive_learning_b200 import _lib
        for kw, knobs in (({"pipelined": True}, {"SCL_KNN_GROUP_M": 1}), ({"two_phase": False}, {})):
            with _lib.tuning(**knobs):
                other = retrieval.ShardedKDTree.__new__(retrieval.ShardedKDTree)
                other.__dict__.update(tree.__dict__)
                other.two_phase, other.pipelined = kw.get("two_phase", True), kw.get("pipelined", False)
                d3, i3 = other.query_device(torch.tensor(qry, device="cuda"), k)
                assert torch.equal(i, i3) and torch.equal(d, d3), kw
        if rank == 0:
            np.savez(out, d=d.cpu().numpy(), i=i.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_sharded_kdtree_two_processes_nccl(cuda_lib, tmp_path):
    """one process per GPU, database rows split over the ranks, ONE packed NCCL all-gather, merge kernel --
    the exact global top-k on every rank, bit-identical to float64 brute force (ties across shards ordered by index)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "nccl.npz")
    mp.spawn(_nccl_worker, args=(2, port, out), nprocs=2, join=True)
    db, qry, info = synth.trajectory_problem(R=40000, Q=300, D=128, seed=9, stop_frac=0.05, stop_len=(100, 200))
    db = np.concatenate([db, db[:2000]], 0)
    rd, ri = orr.knn_bruteforce_exact(db, qry, 25)
    got = np.load(out)
    check_exact(got["d"], got["i"], rd, ri)


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
