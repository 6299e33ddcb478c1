"""Tensor-kernel time of ONE launch over 10 000 queries as a function of the shard size: fixed cost vs per-row cost."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import retrieval, _lib
Q, D, k = 10000, 4096, 25
g = torch.Generator(device="cuda").manual_seed(42)
full = torch.randn((1000000, D), generator=g, device="cuda")
qry = full[torch.randint(0, 31250, (Q,), generator=g, device="cuda")] + 0.5 * torch.randn((Q, D), generator=g, device="cuda")
ub = torch.empty((Q, k), dtype=torch.float32, device="cuda")
for R in (31250, 62500, 125000, 250000, 500000, 1000000):
    tree = retrieval.KDTree(full[:R])
    for _ in range(2):
        st = tree.query_launch(qry, k); tree.query_begin_group(st, qry, k, -1, ub)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        st = tree.query_launch(qry, k)
        e1.record()
        tree.query_begin_group(st, qry, k, -1, ub)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    print(f"R={R:8d}  prep+tensor {ms:8.3f} ms  {2.0*Q*R*D/(ms*1e-3)/1e12:6.0f} TF/s   ms per 125k rows {ms*125000/R:7.3f}", flush=True)
    del tree
