#!/usr/bin/env python
"""One small call per kernel family, for compute-sanitizer (tools/sanitize.sh).  Results are checked against nothing
here -- the parity tests do that -- the point is that every kernel family executes under memcheck / racecheck."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import _lib, losses, netvlad, retrieval, synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"


def wms():
    emb, dist, _ = synth.wms_batch(T=3, P=12, N=12, D=512, seed=1)
    for knobs in ({"SCL_WMS_STREAM": 1, "SCL_WMS_STREAM_CFG": 6}, {"SCL_WMS_STREAM": 1, "SCL_WMS_STREAM_CFG": 2},
                  {"SCL_WMS_STREAM": 0}, {"SCL_WMS_STREAM": 0, "SCL_WMS_CHUNKED": 1}):
        with _lib.tuning(**knobs):
            losses.wms_loss_value_and_grad(dist, emb, 0.8, 15.0)


def tuples():
    rng = np.random.default_rng(0)
    emb = synth.tuple_descriptors(rng, 4, 5, 6, 256, other=True)
    losses.tuple_loss_value_and_grad("quadruplet_loss", emb.reshape(-1, 256), 4, 5, 6, m1=0.1, m2=0.2)
    losses.tuple_loss_value_and_grad("lazy_triplet_loss", emb[:, :12].reshape(-1, 256), 4, 5, 6, m1=0.1)


def flat():
    rng = np.random.default_rng(0)
    emb = rng.standard_normal((256, 256)).astype(np.float32)
    losses.ms_loss_value_and_grad(losses.ms_labels(8, 15, 16), emb)


def nv():
    x, aw, cc, V, m, var = synth.netvlad_problem(B=2, H=12, W=11, Dout=128, seed=1)
    xt = torch.tensor(x, device="cuda", requires_grad=True)
    wt = torch.tensor(aw, device="cuda", requires_grad=True)
    ct = torch.tensor(cc, device="cuda", requires_grad=True)
    v = netvlad.netVLAD(xt, wt, ct)
    y = netvlad.pca_project(v, V, m, var)
    y.sum().backward()


def knn():
    db, qry, info = synth.trajectory_problem(R=5000, Q=300, D=64, seed=3, stop_frac=0.2, stop_len=(100, 200))
    tree = retrieval.KDTree(db)
    with _lib.tuning(SCL_KNN_CHUNK_Q=256):
        d, i = tree.query(qry, k=25, force_path=2)
    print("knn stats", tree.stats())
    tree.query(qry[:8], k=25, force_path=1)
    for v in (1, 3):
        with _lib.tuning(SCL_KNN_TC_VARIANT=v):
            tree.query(qry[:130], k=5, force_path=2)
    retrieval.recall_at_n(np.abs(np.random.default_rng(0).standard_normal((20, 25))) * 10)
    # sharded two-phase protocol on one GPU: 2 shards, tensor launch with group signals, second stream per query group
    q = torch.tensor(qry, device="cuda")
    trees = [retrieval.KDTree(db[:2500]), retrieval.KDTree(db[2500:], index_offset=2500)]
    with _lib.tuning(SCL_KNN_GROUP_M=1):
        n_groups, gq = retrieval.KDTree.query_groups(64, 300)
        side = torch.cuda.Stream()
        states = [t.query_launch(q, 25) for t in trees]
        dd = torch.empty((300, 25), dtype=torch.float64, device="cuda")
        ii = torch.empty((300, 25), dtype=torch.int64, device="cuda")
        with torch.cuda.stream(side):
            for g in range(n_groups):
                q0, nq = g * gq, min(gq, 300 - g * gq)
                ub = torch.empty((2, nq, 25), dtype=torch.float32, device="cuda")
                for r in range(2):
                    trees[r].query_begin_group(states[r], q, 25, g, ub[r])
                bound = retrieval.bound_reduce(ub)
                packed = torch.empty((2, 2, nq, 25), dtype=torch.int64, device="cuda")
                for r in range(2):
                    trees[r].query_end_group(states[r], q, 25, g, bound, (packed[r, 0].view(torch.float64), packed[r, 1]))
                retrieval.topk_merge_packed(packed, 2, nq, 25, out=(dd[q0:q0 + nq], ii[q0:q0 + nq]))
        torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    assert np.array_equal(ii.cpu().numpy(), i), "two-phase != single index"
    print("knn two-phase groups", n_groups, trees[0].stats())


cases = {"wms": wms, "tuples": tuples, "flat": flat, "netvlad": nv, "knn": knn}
for name, fn in cases.items():
    if which in ("all", name):
        fn()
        torch.cuda.synchronize()
        print("ok", name, flush=True)
