"""Tensor-pass time of scl_knn_query at small shards (what a rank sees at N = 8): rows x knobs."""
import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import retrieval, _lib
import ctypes as C
R = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
Q, D = 10000, 4096
g = torch.Generator(device="cuda").manual_seed(42)
db = torch.randn((R, D), generator=g, device="cuda")
qry = db[torch.randint(0, R, (Q,), generator=g, device="cuda")] + 0.5 * torch.randn((Q, D), generator=g, device="cuda")
tree = retrieval.KDTree(db)
L = _lib.lib()
def run(tag, **knobs):
    with _lib.tuning(**knobs):
        for _ in range(2): tree.query_device(qry, k=25)
        torch.cuda.synchronize()
        L.scl_knn_timing(1, None, None)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): tree.query_device(qry, k=25)
        e1.record(); torch.cuda.synchronize()
        ms, n = C.c_double(), C.c_int()
        L.scl_knn_timing(0, C.byref(ms), C.byref(n))
        st = tree.stats()
        fl = 2.0 * Q * R * D
        print(f"R={R} {tag:28s} step {e0.elapsed_time(e1)/5:7.3f} ms  tensor {ms.value/max(n.value,1):7.3f} ms = {fl/(ms.value/max(n.value,1)*1e-3)/1e12:6.0f} TF/s  chunks {st['chunks']} cert {st['n_certified']}", flush=True)
run("default")
run("one chunk", SCL_KNN_CHUNK_Q=0)
run("no pacing", SCL_KNN_SYNC=0)
run("one chunk, no pacing", SCL_KNN_CHUNK_Q=0, SCL_KNN_SYNC=0)
for nr in (2, 4, 6, 16, 24):
    run(f"one chunk, ranges={nr}", SCL_KNN_CHUNK_Q=0, SCL_KNN_RANGES=nr)
