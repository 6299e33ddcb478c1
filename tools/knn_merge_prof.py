"""One rank's first phase at an N = 8 shard (R = 125 000): the target of an ncu capture of knn_cand_merge_kernel."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import retrieval
R, Q, D, k = 125000, 10000, 4096, 25
g = torch.Generator(device="cuda").manual_seed(42)
db = torch.randn((R, D), generator=g, device="cuda")
qry = db[torch.randint(0, R, (Q,), generator=g, device="cuda")] + 0.5 * torch.randn((Q, D), generator=g, device="cuda")
tree = retrieval.KDTree(db)
ub = torch.empty((1, Q, k), dtype=torch.float32, device="cuda")
out = (torch.empty((Q, k), dtype=torch.float64, device="cuda"), torch.empty((Q, k), dtype=torch.int64, device="cuda"))
for _ in range(3):
    st = tree.query_begin(qry, k, ub[0])
    tree.query_end(st, qry, k, retrieval.bound_reduce(ub), out)
torch.cuda.synchronize()
print(tree.stats())
