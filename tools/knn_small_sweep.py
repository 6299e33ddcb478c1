"""Tensor-kernel time of ONE launch over 10 000 queries at an N = 8 shard (125 000 rows): knob sweep."""
import os, sys, torch, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import retrieval, _lib
R = int(sys.argv[1]) if len(sys.argv) > 1 else 125000
Q, D, k = 10000, 4096, 25
g = torch.Generator(device="cuda").manual_seed(42)
db = torch.randn((R, D), generator=g, device="cuda")
qry = db[torch.randint(0, R, (Q,), generator=g, device="cuda")] + 0.5 * torch.randn((Q, D), generator=g, device="cuda")
tree = retrieval.KDTree(db)
L = _lib.lib()
ub = torch.empty((Q, k), dtype=torch.float32, device="cuda")
def run(tag, **knobs):
    with _lib.tuning(**knobs):
        for _ in range(2):
            st = tree.query_launch(qry, k); tree.query_begin_group(st, qry, k, -1, ub)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ts = []
        for _ in range(5):
            e0.record()
            st = tree.query_launch(qry, k)
            e1.record()
            tree.query_begin_group(st, qry, k, -1, ub)
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        print(f"R={R} {tag:40s} prep+tensor {ms:7.3f} ms = {2.0*Q*R*D/(ms*1e-3)/1e12:6.0f} TF/s", flush=True)
run("default (variant 2)")
run("variant 3 (256x512 tiles)", SCL_KNN_TC_VARIANT=3)
for gm in (10, 20, 40):
    run(f"variant 3, group_m={gm}", SCL_KNN_TC_VARIANT=3, SCL_KNN_GROUP_M=gm)
for sb in (1, 2, 8):
    run(f"variant 3, subs={sb}", SCL_KNN_TC_VARIANT=3, SCL_KNN_SYNC_SUBS=sb)
for w in (2, 8):
    run(f"variant 3, window={w}", SCL_KNN_TC_VARIANT=3, SCL_KNN_SYNC_WINDOW=w)
for nr in (18, 25, 37, 49):
    run(f"variant 3, ranges={nr}", SCL_KNN_TC_VARIANT=3, SCL_KNN_RANGES=nr)
