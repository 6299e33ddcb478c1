"""What ONE rank of an N-way sharded retrieval spends per step, measured on one GPU: the G shards of the config-4 database
are all resident (16 GB + 8 GB of shadows), each runs scl_knn_query_begin, the bounds are MIN-reduced in place of the
all-reduce, and scl_knn_query_end + the packed merge are timed for shard 0.  Compared with the plain protocol."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from soft_contrastive_learning_b200 import retrieval, _lib
G = int(sys.argv[1]) if len(sys.argv) > 1 else 8
R, Q, D, k = 1000000, 10000, 4096, 25
g = torch.Generator(device="cuda").manual_seed(42)
per = R // G
trees = []
for r in range(G):
    trees.append(retrieval.KDTree(torch.randn((per, D), generator=g, device="cuda"), index_offset=r * per))
src = torch.randint(0, R, (Q,), generator=g, device="cuda")
qry = torch.stack([trees[int(s) // per].db[int(s) % per] for s in src.tolist()]) + 0.5 * torch.randn((Q, D), generator=g, device="cuda")

def ev():
    return torch.cuda.Event(enable_timing=True)

def run(tag, **knobs):
    with _lib.tuning(**knobs):
        packed = torch.empty((G, 2, Q, k), dtype=torch.int64, device="cuda")
        plain = torch.empty((G, 2, Q, k), dtype=torch.int64, device="cuda")
        for r in range(G):
            trees[r].query_device(qry, k, out=(plain[r, 0].view(torch.float64), plain[r, 1]))
        dref, iref = retrieval.topk_merge_packed(plain, G, Q, k)
        bnds = [torch.empty((Q, k), dtype=torch.float32, device="cuda") for _ in range(G)]
        states = [trees[r].query_begin(qry, k, bnds[r]) for r in range(G)]
        bound = retrieval.bound_reduce(torch.stack(bnds))
        for r in range(G):
            trees[r].query_end(states[r], qry, k, bound, (packed[r, 0].view(torch.float64), packed[r, 1]))
        d, i = retrieval.topk_merge_packed(packed, G, Q, k)
        same = bool(torch.equal(i, iref) and torch.equal(d, dref))
        real = int((packed[:, 1] >= 0).sum())
        st = trees[0].stats()
        # timing of rank 0's share
        t = {"plain": 0.0, "begin": 0.0, "end": 0.0, "merge": 0.0}
        n = 5
        for it in range(n + 1):
            e = [ev() for _ in range(5)]
            e[0].record()
            trees[0].query_device(qry, k, out=(plain[0, 0].view(torch.float64), plain[0, 1]))
            e[1].record()
            s0 = trees[0].query_begin(qry, k, bnds[0])
            e[2].record()
            trees[0].query_end(s0, qry, k, bound, (packed[0, 0].view(torch.float64), packed[0, 1]))
            e[3].record()
            retrieval.topk_merge_packed(packed, G, Q, k)
            e[4].record()
            torch.cuda.synchronize()
            if it:
                for j, key in enumerate(t):
                    t[key] += e[j].elapsed_time(e[j + 1]) / n
        print(f"G={G} {tag:12s} same={same} rows returned {real} (plain {G*Q*k}) rank0 {st}  "
              f"plain {t['plain']:.3f} ms | begin {t['begin']:.3f} + end {t['end']:.3f} = {t['begin']+t['end']:.3f} ms, merge {t['merge']:.3f}", flush=True)

run("default")
run("one chunk", SCL_KNN_CHUNK_Q=0)
