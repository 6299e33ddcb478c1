"""Re-entrancy of the C ABI (SURVEY.md 8b "threading"): the reference drives one session from up to three Python
threads (train/train.py:286 training, :215 evaluation loss, :152-158 localization).  Three host threads call the loss,
NetVLAD+PCA and retrieval entry points concurrently, each on its own CUDA stream; every result must equal the one the
same call gives alone."""
import threading

import numpy as np
import pytest
import torch

from soft_contrastive_learning_b200 import synth

pytestmark = pytest.mark.gpu


def test_three_host_threads_share_the_library(cuda_lib):
    from soft_contrastive_learning_b200 import losses, netvlad, retrieval
    emb, dist, _ = synth.wms_batch(T=160, P=12, N=12, D=512, seed=5)
    e, d = torch.tensor(emb, device="cuda"), torch.tensor(dist, device="cuda")
    p = losses._ms_params(0.8, 15.0)
    x, aw, cc, V, m, var = synth.netvlad_problem(B=3, H=5, W=6, C=512, K=64, Dout=128, seed=8)
    xt, awt, cct = (torch.tensor(a, device="cuda") for a in (x, aw, cc))
    Vt, mt, vart = (torch.tensor(a, device="cuda") for a in (V, m, var))
    db, qry, _, _, _ = synth.retrieval_problem(R=20000, Q=256, D=256, seed=6)
    tree = retrieval.KDTree(db)
    qt = torch.tensor(qry, device="cuda")

    def run_loss():
        l, g, _, _ = losses._wms_tuple_raw(e, d, p, need_grad=True)
        return l.clone(), g.clone()

    def run_head():
        with torch.no_grad():
            return netvlad.pca_project(netvlad.netVLAD(xt, awt, cct), Vt, mt, vart).clone()

    def run_knn():
        dd, ii = tree.query_device(qt, k=25)
        return dd.clone(), ii.clone()

    alone = [run_loss(), run_head(), run_knn()]
    torch.cuda.synchronize()
    results, errors = {}, []

    def worker(name, fn, reps):
        try:
            with torch.cuda.stream(torch.cuda.Stream()):
                out = None
                for _ in range(reps):
                    out = fn()
                torch.cuda.current_stream().synchronize()
                results[name] = out
        except Exception as exc:                      # surfaced below: a thread must not die silently
            errors.append((name, exc))

    threads = [threading.Thread(target=worker, args=a) for a in (("loss", run_loss, 20), ("head", run_head, 20), ("knn", run_knn, 5))]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    assert torch.equal(results["loss"][0], alone[0][0]) and torch.equal(results["loss"][1], alone[0][1])
    assert torch.equal(results["head"], alone[1])
    assert torch.equal(results["knn"][0], alone[2][0]) and torch.equal(results["knn"][1], alone[2][1])
