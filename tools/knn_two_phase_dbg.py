import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import retrieval, synth
R, Q, D, k, G = 40000, 300, 256, 25, 4
db, qry, *_ = synth.retrieval_problem(R=R, Q=Q, D=D, seed=15)
q = torch.tensor(qry, device="cuda")
trees, states, bnds = [], [], []
for r in range(G):
    lo, hi = retrieval.shard_bounds(R, G, r)
    trees.append(retrieval.KDTree(db[lo:hi], index_offset=lo))
    b = torch.empty((Q, k), dtype=torch.float32, device="cuda")
    states.append(trees[-1].query_begin(q, k, b)); bnds.append(b)
B = torch.stack(bnds)
bound = retrieval.bound_reduce(B)
qn2 = (q.double() ** 2).sum(1)
dbt = torch.tensor(db, device="cuda").double()
d2 = (qn2[:, None] + (dbt ** 2).sum(1)[None] - 2 * q.double() @ dbt.T)     # [Q,R]
gk = d2.sort(1).values[:, k - 1] - qn2
print("global k-th exact score  :", gk[:6].cpu().numpy())
print("reduced bound            :", bound[:6].cpu().numpy())
for r in range(G):
    lo, hi = retrieval.shard_bounds(R, G, r)
    loc = d2[:, lo:hi].sort(1).values - qn2[:, None]
    print(f"shard {r}: local k-th", loc[:3, k - 1].cpu().numpy(), "bound", B[r, :3, k - 1].cpu().numpy(), "local 64th", loc[:3, 63].cpu().numpy(),
          "rows <= reduced bound (mean)", float((loc <= bound[:, None].double()).sum(1).float().mean()))
    out = (torch.empty((Q, k), dtype=torch.float64, device="cuda"), torch.empty((Q, k), dtype=torch.int64, device="cuda"))
    trees[r].query_end(states[r], q, k, bound, out)
    print("   stats", trees[r].stats(), "returned", int((out[1] >= 0).sum()))
