"""What ONE rank of an N-way sharded retrieval spends per step, measured on one GPU: the G shards of the config-4 database
are all resident (16 GB + 8 GB of shadows); the other ranks' score bounds and packed lists are computed once and stand in
for the all-gathers; rank 0's step is timed in three forms: plain (full local top-k), two-phase on one stream, two-phase
pipelined over the query groups of the tensor launch (second stream)."""
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

# reference: plain protocol over all shards; the other ranks' phase-1 / phase-2 messages
plain = torch.empty((G, 2, Q, k), dtype=torch.int64, device="cuda")
for r in range(G):
    trees[r].query_device(qry, k, out=(plain[r, 0].view(torch.float64), plain[r, 1]))
dref, iref = retrieval.topk_merge_packed(plain, G, Q, k)
ub_all = torch.empty((G, Q, k), dtype=torch.float32, device="cuda")
states = [trees[r].query_begin(qry, k, ub_all[r]) for r in range(G)]
bound = retrieval.bound_reduce(ub_all)
packed = torch.empty((G, 2, Q, k), dtype=torch.int64, device="cuda")
for r in range(G):
    trees[r].query_end(states[r], qry, k, bound, (packed[r, 0].view(torch.float64), packed[r, 1]))
d, i = retrieval.topk_merge_packed(packed, G, Q, k)
print(f"G={G} two-phase == plain: {bool(torch.equal(i, iref) and torch.equal(d, dref))}; rows returned {int((packed[:, 1] >= 0).sum())} (plain {G*Q*k})", flush=True)
del states

t0 = trees[0]
n_groups, gq = retrieval.KDTree.query_groups(D, Q)
side = torch.cuda.Stream()
dout = torch.empty((Q, k), dtype=torch.float64, device="cuda")
iout = torch.empty((Q, k), dtype=torch.int64, device="cuda")

def step_plain():
    t0.query_device(qry, k, out=(plain[0, 0].view(torch.float64), plain[0, 1]))
    retrieval.topk_merge_packed(plain, G, Q, k, out=(dout, iout))

def step_two_phase():
    st = t0.query_begin(qry, k, ub_all[0])
    b = retrieval.bound_reduce(ub_all)
    t0.query_end(st, qry, k, b, (packed[0, 0].view(torch.float64), packed[0, 1]))
    retrieval.topk_merge_packed(packed, G, Q, k, out=(dout, iout))

gbuf = []
for gi in range(n_groups):
    q0, nq = gi * gq, min(gq, Q - gi * gq)
    gbuf.append((ub_all[:, q0:q0 + nq].contiguous(), packed[:, :, q0:q0 + nq].contiguous()))

def step_pipelined():
    main = torch.cuda.current_stream()
    st = t0.query_launch(qry, k)
    with torch.cuda.stream(side):
        for gi in range(n_groups):
            q0, nq = gi * gq, min(gq, Q - gi * gq)
            ubg, pkg = gbuf[gi]
            t0.query_begin_group(st, qry, k, gi, ubg[0])
            b = retrieval.bound_reduce(ubg)
            t0.query_end_group(st, qry, k, gi, b, (pkg[0, 0].view(torch.float64), pkg[0, 1]))
            retrieval.topk_merge_packed(pkg, G, nq, k, out=(dout[q0:q0 + nq], iout[q0:q0 + nq]))
    main.wait_stream(side)

L = _lib.lib()
import ctypes as C
for name, fn in (("plain", step_plain), ("two-phase", step_two_phase), (f"pipelined x{n_groups}", step_pipelined)):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ok = bool(torch.equal(iout, iref) and torch.equal(dout, dref))
    L.scl_knn_timing(1, None, None)
    e0, e1 = ev(), ev()
    n = 8
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms, nc = C.c_double(), C.c_int()
    L.scl_knn_timing(0, C.byref(ms), C.byref(nc))
    print(f"G={G} rank 0 step, {name:14s}: {e0.elapsed_time(e1)/n:7.3f} ms (tensor kernel {ms.value/max(nc.value,1):6.3f} ms) exact={ok} {t0.stats()}", flush=True)

# time-line of one pipelined step: event offsets (ms) from the start of the step
def timeline():
    main = torch.cuda.current_stream()
    marks = []
    def mark(name, stream):
        e = ev(); e.record(stream); marks.append((name, e))
    mark("start", main)
    st = t0.query_launch(qry, k)
    mark("tensor kernel done (main)", main)
    with torch.cuda.stream(side):
        for gi in range(n_groups):
            q0, nq = gi * gq, min(gq, Q - gi * gq)
            ubg, pkg = gbuf[gi]
            t0.query_begin_group(st, qry, k, gi, ubg[0])
            mark(f"g{gi} candidates merged", side)
            b = retrieval.bound_reduce(ubg)
            mark(f"g{gi} bound", side)
            t0.query_end_group(st, qry, k, gi, b, (pkg[0, 0].view(torch.float64), pkg[0, 1]))
            mark(f"g{gi} end_group returned", side)
            retrieval.topk_merge_packed(pkg, G, nq, k, out=(dout[q0:q0 + nq], iout[q0:q0 + nq]))
            mark(f"g{gi} shard merge", side)
    main.wait_stream(side)
    mark("step done (main)", main)
    torch.cuda.synchronize()
    return [(n, marks[0][1].elapsed_time(e)) for n, e in marks]
for _ in range(2):
    tl = timeline()
print("time-line of a pipelined step (ms from start):")
for n, t in tl:
    print(f"  {t:8.3f}  {n}")

def probe(extra):
    main = torch.cuda.current_stream()
    marks = []
    def mark(name, stream):
        e = ev(); e.record(stream); marks.append((name, e))
    small = torch.zeros(1024, device="cuda")
    mark("start", main)
    st = t0.query_launch(qry, k)
    mark("tensor kernel done (main)", main)
    with torch.cuda.stream(side):
        for gi in range(n_groups):
            ubg, pkg = gbuf[gi]
            t0.query_begin_group(st, qry, k, gi, ubg[0])
            mark(f"g{gi} candidates merged", side)
            if extra:
                small.add_(1.0)
                mark(f"g{gi} trivial kernel", side)
                b = retrieval.bound_reduce(ubg)
                mark(f"g{gi} bound", side)
    main.wait_stream(side)
    mark("step done (main)", main)
    torch.cuda.synchronize()
    return [(n, marks[0][1].elapsed_time(e)) for n, e in marks]
for gm_knob in (None, 10, 5):
    _lib.set_tuning("SCL_KNN_GROUP_M", gm_knob)
    n_groups, gq = retrieval.KDTree.query_groups(D, Q)
    gbuf = []
    for gi in range(n_groups):
        q0, nq = gi * gq, min(gq, Q - gi * gq)
        gbuf.append((ub_all[:, q0:q0 + nq].contiguous(), packed[:, :, q0:q0 + nq].contiguous()))
    for _ in range(2):
        tl = probe(False)
    print(f"probe (group_m={gm_knob}, {n_groups} groups of {gq}):")
    for n, t in tl:
        print(f"  {t:8.3f}  {n}")
_lib.set_tuning("SCL_KNN_GROUP_M", None)
