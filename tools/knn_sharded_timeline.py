"""Time-line of one sharded two-phase retrieval step on N GPUs (torchrun): where the time between the end of the tensor
kernel and the end of the step goes (candidate merge, all-gathers, bound, rescore, host sync, shard merge).
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/knn_sharded_timeline.py"""
import os, sys, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import retrieval
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
dist.init_process_group("nccl")
R, Q, D, k = 1000000, 10000, 4096, 25
per = R // world
g = torch.Generator(device="cuda").manual_seed(100 + rank)
db = torch.randn((per, D), generator=g, device="cuda")
gq = torch.Generator(device="cuda").manual_seed(7)
qry = torch.randn((Q, D), generator=gq, device="cuda")          # same on every rank
tree = retrieval.KDTree(db, index_offset=rank * per)
G = world

def ev():
    return torch.cuda.Event(enable_timing=True)

def step(trace):
    marks = []
    def mark(name):
        if trace:
            e = ev(); e.record(); marks.append((name, e))
    mark("start")
    ub_all = torch.empty((G, Q, k), dtype=torch.float32, device="cuda")
    mine = torch.empty((2, Q, k), dtype=torch.int64, device="cuda")
    packed = torch.empty((G, 2, Q, k), dtype=torch.int64, device="cuda")
    st = tree.query_launch(qry, k)
    mark("prep + tensor kernel")
    tree.query_begin_group(st, qry, k, -1, ub_all[rank])
    mark("candidate merge")
    dist.all_gather_into_tensor(ub_all.view(G * Q, k), ub_all[rank])
    mark("all-gather of the bounds (1 MB per rank)")
    bound = retrieval.bound_reduce(ub_all)
    mark("bound reduce")
    tree.query_end_group(st, qry, k, -1, bound, (mine[0].view(torch.float64), mine[1]))
    mark("cutoff + rescore + certificate + host sync")
    dist.all_gather_into_tensor(packed.view(G * 2 * Q, k), mine.view(2 * Q, k))
    mark("all-gather of the packed lists (4 MB per rank)")
    d, i = retrieval.topk_merge_packed(packed, G, Q, k)
    mark("shard merge")
    return marks

for _ in range(3):
    step(False)
torch.cuda.synchronize(); dist.barrier()
acc = None
n = 8
for _ in range(n):
    m = step(True)
    torch.cuda.synchronize()
    t = [m[j][1].elapsed_time(m[j + 1][1]) for j in range(len(m) - 1)]
    acc = t if acc is None else [a + b for a, b in zip(acc, t)]
names = [x[0] for x in m[1:]]
acc = torch.tensor([a / n for a in acc], device="cuda")
allr = [torch.empty_like(acc) for _ in range(world)]
dist.all_gather(allr, acc)
if rank == 0:
    tab = torch.stack(allr).cpu()
    print(f"N = {world}: mean over {n} steps, ms (rank 0 | min over ranks | max over ranks)")
    for j, nm in enumerate(names):
        print(f"  {tab[0, j]:7.3f} | {tab[:, j].min():7.3f} | {tab[:, j].max():7.3f}   {nm}")
    print(f"  {tab[0].sum():7.3f}   total on rank 0")
dist.destroy_process_group()
