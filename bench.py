#!/usr/bin/env python
"""bench.py -- headline benchmark of the descriptor-space hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload retrieval|wms]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Primary line (metric A): top-25 queries/s against a 1M x 4096 fp32 database (BASELINE config 4: 10 000 queries per
step).  With N GPUs the 1M rows are split contiguously over the ranks (strong scaling), queries are replicated, the
per-shard lists are all-gathered over NCCL and merged; `value` = queries / (max-over-ranks device time).
Secondary (metric B, N=1 only): fused wms forward+backward tuples/s (S=25, D=4096).

One JSON line on stdout (rank 0).  `value` is timed with inputs resident in HBM; `e2e` is the same call fed from
pinned HOST buffers with the host<->device copies inside the timed region.  `roofline` describes the dominant
kernel (tcgen05 distance GEMM) with its own device time measured by CUDA events on the launching stream.
`cpu_baseline` / `--impl reference` time the reference's own CPU call -- sklearn KDTree.query exactly as
evaluation/top-n.py:103-106 -- on a bounded sample and scale it linearly in the number of rows (a 4096-d KD-tree
degenerates to a brute-force scan, so linear scaling in R is generous to it); the wms baseline is the oracle port.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_FULL, D_FULL, K_TOP = 1_000_000, 4096, 25
Q_STEP = 10_000


def workload_config(rows, queries, shards):
    """The workload both arms (ours and --impl reference) are measured on: identical dict, identical wording."""
    return {"workload": f"top-{K_TOP} exact retrieval, {queries} queries/step vs {rows}x{D_FULL} fp32 db (BASELINE config 4)",
            "rows": rows, "queries_per_step": queries, "dim": D_FULL, "k": K_TOP, "db_shards": shards}


def profile_traffic(fname):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, parsed from the committed `ncu --set full` digest
    under profiles/ (tools/ncu_digest.py writes 'dram read  <x> Mbyte|Gbyte').  None when the digest is absent."""
    import re
    p = os.path.join(ROOT, "profiles", fname)
    if not os.path.exists(p):
        return None
    tot, seen = 0.0, 0
    for ln in open(p):
        m = re.match(r"\s*dram (read|write)\s+([0-9.]+)\s+([KMG]?)byte", ln)
        if m:
            tot += float(m.group(2)) * {"": 1.0, "K": 1e3, "M": 1e6, "G": 1e9}[m.group(3)]
            seen += 1
    return tot if seen >= 2 else None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "tf_burst": j["bf16_tflops"], "tf_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]),
                "source": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:]      # upper half = samples taken under load
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# reference arm / CPU baselines
# ------------------------------------------------------------------------------------------------
def kdtree_reference(steps, warmup, R_s=10_000, Q_s=None, D=D_FULL, k=K_TOP):
    """The reference's own call (evaluation/top-n.py:103-106) on a bounded sample, scaled linearly to R_FULL rows.
    The reference script issues it once, single-threaded; KDTree.query releases the GIL, so the same call over query
    chunks from a thread pool uses every host core -- that is the number reported (the single-threaded one is kept in
    the sample text)."""
    from concurrent.futures import ThreadPoolExecutor
    from sklearn.neighbors import KDTree
    cores = os.cpu_count() or 1
    Q_s = Q_s or 8 * cores
    rng = np.random.default_rng(42)
    ref = rng.standard_normal((R_s, D), dtype=np.float32)
    qry = ref[rng.integers(0, R_s, Q_s)] + 0.5 * rng.standard_normal((Q_s, D), dtype=np.float32)
    t0 = time.perf_counter()
    tree = KDTree(ref)
    build_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    tree.query(qry[:32], k=k, return_distance=True, sort_results=True)
    qps_single = 32 / (time.perf_counter() - t0)
    chunks = [c for c in np.array_split(np.arange(Q_s), cores) if len(c)]
    times = []
    with ThreadPoolExecutor(cores) as ex:
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            list(ex.map(lambda ix: tree.query(qry[ix], k=k, return_distance=True, sort_results=True), chunks))
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    t = float(np.mean(times))
    qps_sample = Q_s / t
    # the stronger CPU number: float32 sgemm shortlist on every host core + float64 rescore (the oracle port)
    from oracle import retrieval as orr
    Rb, Qb = 50_000, 256
    refb = rng.standard_normal((Rb, D), dtype=np.float32)
    qb = refb[rng.integers(0, Rb, Qb)] + 0.5 * rng.standard_normal((Qb, D), dtype=np.float32)
    t0 = time.perf_counter()
    orr.knn_sgemm_allcores(refb, qb, k)
    tb = time.perf_counter() - t0
    return {"value": qps_sample * R_s / R_FULL, "unit": "queries/s", "cores": cores, "kind": "reference",
            "all_core_bruteforce": {"value": Qb / tb * Rb / R_FULL, "unit": "queries/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"NumPy sgemm shortlist + float64 rescore on [{Rb}x{D}] x q[{Qb}], scaled by {Rb}/{R_FULL} rows"},
            "sample": f"sklearn KDTree(ref[{R_s}x{D}]).query(q[{Q_s}], k={k}, sort_results=True) over {cores} threads: "
                      f"{qps_sample:.2f} q/s measured ({qps_single:.2f} q/s single-threaded, as the reference script calls it), "
                      f"scaled by {R_s}/{R_FULL} rows (linear in R); tree build {build_s:.1f} s not counted",
            "ms_per_step": t * 1e3}


def wms_cpu_port(T=32, S=25, D=D_FULL, reps=3):
    """Oracle port (float64 torch-CPU autograd transcription of model/losses.py:5-60) on all host threads."""
    import torch
    from oracle import losses as ol
    from soft_contrastive_learning_b200 import synth
    torch.set_num_threads(os.cpu_count() or 1)
    emb, dist, _ = synth.wms_batch(T=T, P=12, N=12, D=D, seed=42)
    e64, d64 = emb.astype(np.float64), torch.as_tensor(dist.astype(np.float64))
    ol.value_and_grad(lambda e: ol.wms_loss_tuples(d64, e, 0.8, 15.0), [e64])
    t0 = time.perf_counter()
    for _ in range(reps):
        ol.value_and_grad(lambda e: ol.wms_loss_tuples(d64, e, 0.8, 15.0), [e64])
    t = (time.perf_counter() - t0) / reps
    return {"value": T / t, "unit": "tuples/s", "cores": os.cpu_count() or 1, "kind": "port",
            "sample": f"oracle (torch-CPU float64 autograd transcription of losses.py:5-60) fwd+bwd on {T} tuples x {S} x {D}, mean of {reps}"}


def run_reference(args, out):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "wms":
        cb = wms_cpu_port(reps=max(1, args.steps))
        line = {"impl": "reference", "metric": "wms loss fwd+bwd tuples/s", "value": cb["value"], "unit": "tuples/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 32 / cb["value"] * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "wms tuple mode T=32 S=25 D=4096 (BASELINE config 1)"}, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "tuples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    else:
        cb = kdtree_reference(args.steps, args.warmup)
        line = {"impl": "reference", "metric": "top-25 queries/s vs 1Mx4096 db", "value": cb["value"], "unit": "queries/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb.pop("ms_per_step"),
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(R_FULL, Q_STEP, args.gpus),
                "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=out, flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def timed(torch, fn, steps, warmup, dist_mod=None):
    for _ in range(warmup):
        fn()
    if dist_mod is not None:
        dist_mod.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist_mod is not None:
        t = torch.tensor([ms], device="cuda")
        dist_mod.all_reduce(t, op=dist_mod.ReduceOp.MAX)
        ms = float(t.item())
        dist_mod.barrier()
    return ms / steps


def make_database(torch, synth, data, lo, hi, R, D, Q, rank, dist_mod):
    """This rank's rows [lo, hi) of the synthetic database and the (replicated) queries, generated on the device.
    data = "gaussian": iid N(0,1) rows, queries = perturbed random rows (planted neighbours).
    data = "clustered": synth.trajectory_* -- consecutive frames of one drive (near-duplicates) with stops where hundreds
    of frames coincide up to sensor noise; queries = frames of a second drive (perturbed database frames)."""
    n = hi - lo
    db = torch.empty((n, D), dtype=torch.float32, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(42 + rank)
    chunk = 65536
    info = {}
    if data == "gaussian":
        for r0 in range(0, n, chunk):          # chunked: randn's temporaries stay small
            r1 = min(n, r0 + chunk)
            db[r0:r1] = torch.randn((r1 - r0, D), generator=g, device="cuda")
        qnoise = 0.5
    else:
        rng = np.random.default_rng(1234)      # the layout of the WHOLE drive is the same on every rank
        s_all, stopped_all = synth.trajectory_layout(R, rng, stop_frac=0.06, stop_len=(100, 300))
        seg_len = 64
        n_anchor = int(s_all[-1] // seg_len) + 2
        ga = torch.Generator(device="cuda").manual_seed(99)
        anchors = torch.randn((n_anchor, D), generator=ga, device="cuda")
        s_t = torch.tensor(s_all[lo:hi], device="cuda")
        st_t = torch.tensor(stopped_all[lo:hi], device="cuda")
        for r0 in range(0, n, chunk):
            r1 = min(n, r0 + chunk)
            noise = torch.randn((r1 - r0, D), generator=g, device="cuda")
            db[r0:r1] = synth.trajectory_rows(s_t[r0:r1], st_t[r0:r1], anchors, seg_len, 0.05, 2e-3, noise)
        del anchors
        qnoise = 0.05
        info["stopped_frac"] = float(stopped_all.mean())
        info["stopped"] = stopped_all
    # queries = perturbed database rows (planted neighbours), identical on every rank
    gq = torch.Generator(device="cuda").manual_seed(7)
    src = torch.randint(0, R, (Q,), generator=gq, device="cuda")
    noise = qnoise * torch.randn((Q, D), generator=gq, device="cuda")
    mine = (src >= lo) & (src < hi)
    qry = torch.zeros((Q, D), dtype=torch.float32, device="cuda")
    qry[mine] = db[(src[mine] - lo)] + noise[mine]
    if dist_mod is not None:
        dist_mod.all_reduce(qry)
    del noise
    if data == "clustered":
        info["stopped_src"] = info.pop("stopped")[src.cpu().numpy()]
        info["queries_on_stops"] = int(info["stopped_src"].sum())
    return db, qry, src, info


def retrieval_run(args, torch, dist_mod, rank, world, R, Q, data="gaussian", steps=None, warmup=None, e2e=True,
                  time_build=False, clocks=True):
    """One retrieval measurement: R database rows split contiguously over the ranks, Q replicated queries per step."""
    from soft_contrastive_learning_b200 import _lib, retrieval, synth
    L = _lib.lib()
    D, k = D_FULL, K_TOP
    steps = steps or args.steps
    warmup = warmup or args.warmup
    lo, hi = retrieval.shard_bounds(R, world, rank)
    db, qry, src, info = make_database(torch, synth, data, lo, hi, R, D, Q, rank, dist_mod)

    # index build from HOST memory (H2D of the shard + shadow build), timed once
    t_build_h2d = None
    if time_build and world == 1:
        host = torch.empty((hi - lo, D), dtype=torch.float32, pin_memory=True)
        host.copy_(db)
        torch.cuda.synchronize()
        del db
        torch.cuda.empty_cache()
        t0 = time.perf_counter()
        db = host.to("cuda", non_blocking=True)
        tree = retrieval.KDTree(db, index_offset=lo)
        torch.cuda.synchronize()
        t_build_h2d = time.perf_counter() - t0
        del host
    else:
        tree = retrieval.KDTree(db, index_offset=lo)
    torch.cuda.synchronize()
    if dist_mod is not None:
        index = retrieval.ShardedKDTree.__new__(retrieval.ShardedKDTree)
        index.group, index.world, index.local, index.two_phase, index._side = None, world, tree, not args.single_phase, None
        index.pipelined = args.group_pipeline
    else:
        index = tree

    out = {}

    def step_dev():
        out["d"], out["i"] = index.query_device(qry, k)

    # correctness guard inside the bench: the planted neighbour must be rank 1 (clustered data: its distance must be the
    # smallest, the frame itself may tie with frames of the same stop)
    step_dev()
    torch.cuda.synchronize()
    if data == "gaussian":
        assert bool((out["i"][:, 0] == src).all()), "planted neighbours not recovered"
    else:
        # frames of one stop tie up to sensor noise, so the planted frame need not be rank 1: hold the first queries that
        # fall on stops (and the first 32 overall) to the exact float64 scan instead, bit for bit
        on_stop = torch.tensor(np.nonzero(info["stopped_src"])[0][:32], device="cuda", dtype=torch.long)
        sel = torch.cat((torch.arange(32, device="cuda"), on_stop))
        d1, i1 = tree.query_device(qry[sel], k, force_path=1)
        assert torch.equal(i1, out["i"][sel]) and torch.equal(d1, out["d"][sel]), "clustered: tensor path != exact scan"
    stats = tree.stats() if data == "gaussian" else None
    if stats is None:
        step_dev()
        torch.cuda.synchronize()
        stats = tree.stats()
    info.pop("stopped_src", None)

    sampler = ClockSampler(torch.cuda.current_device()) if clocks else None
    L.scl_knn_timing(1, None, None)
    if sampler:
        sampler.start()
    ms = timed(torch, step_dev, steps, warmup, dist_mod)
    clk = None
    if sampler:
        # nvidia-smi needs a few hundred ms to deliver samples: when the timed region is shorter than that (small shards at
        # N = 8: 5 steps x 10 ms), the same step keeps running untimed under the sampler until ~0.8 s of load has been seen
        # (ms is the max over ranks, so every rank runs the same number of extra steps and the collectives stay matched)
        loaded_s = ms * 1e-3 * (steps + warmup)
        n_extra = int(np.ceil((0.8 - loaded_s) / (ms * 1e-3))) if loaded_s < 0.8 else 0
        for _ in range(n_extra):
            step_dev()
        torch.cuda.synchronize()
        clk = sampler.stop()
        if n_extra:
            clk["untimed_steps_under_sampler"] = n_extra
    tc_ms, tc_calls = C.c_double(), C.c_int()
    L.scl_knn_timing(0, C.byref(tc_ms), C.byref(tc_calls))
    tc_avg_ms = tc_ms.value / max(1, tc_calls.value)

    ms_e2e = None
    if e2e:
        # end to end: pinned host queries in, host results out, every step
        q_host = torch.empty((Q, D), dtype=torch.float32, pin_memory=True)
        q_host.copy_(qry)
        d_host = torch.empty((Q, k), dtype=torch.float64, pin_memory=True)
        i_host = torch.empty((Q, k), dtype=torch.int64, pin_memory=True)

        def step_e2e():
            if dist_mod is not None:
                d, i = index.query_from_host(q_host, k)      # 1/N of the queries per rank over PCIe, all-gather over NVLink
            else:
                qd = q_host.to("cuda", non_blocking=True)
                d, i = index.query_device(qd, k)
            d_host.copy_(d, non_blocking=True)
            i_host.copy_(i, non_blocking=True)
            torch.cuda.current_stream().synchronize()

        ms_e2e = timed(torch, step_e2e, steps, warmup, dist_mod)
        if data == "gaussian":
            assert np.array_equal(i_host.numpy()[:, 0], src.cpu().numpy())
    rows_local = hi - lo
    del tree, index, db, qry
    torch.cuda.empty_cache()
    flops = 2.0 * Q * rows_local * D
    # launches of this repo's kernels per step: prep + per chunk (tensor pass, candidate merge, rescore, certificate) +
    # stats (+ shard merge); refused queries add the second stage (gather, tensor pass, rescore, select)
    nf = stats["n_fallback"]
    # sharded (two-phase): prep, tensor pass, candidate merge | bound reduce | cutoff, rescore, certificate, stats | shard merge
    launches = 1 + 4 * max(1, stats["chunks"]) + 1 + (1 if world > 1 else 0) + (4 * -(-nf // 2048) if nf else 0)
    if world > 1 and not args.single_phase:
        launches += 2 + (4 * (max(1, stats["chunks"]) - 1) if args.group_pipeline else 0)
    return {"ms": ms, "ms_e2e": ms_e2e, "tc_ms": tc_avg_ms, "flops_per_launch": flops, "stats": stats, "clocks": clk,
            "rows_local": rows_local, "build_s": t_build_h2d, "launches_per_step": launches, "info": info, "steps": steps,
            "warmup": warmup}


def bench_retrieval(args, torch, dist_mod, rank, world, pk):
    R, D, Q, k = args.rows, D_FULL, args.queries, K_TOP
    m = retrieval_run(args, torch, dist_mod, rank, world, R, Q, data=args.data, e2e=True, time_build=args.time_build)
    ms, ms_e2e, tc_avg_ms, stats = m["ms"], m["ms_e2e"], m["tc_ms"], m["stats"]
    achieved = m["flops_per_launch"] / (tc_avg_ms * 1e-3) / 1e12 if tc_avg_ms > 0 else 0.0
    cfg = workload_config(R, Q, world)
    cfg.update({"data_kind": args.data, "rows_per_gpu": m["rows_local"],
                "l2": "inputs larger than L2 (db shard fp32+fp16 >> 126 MB); no flush",
                "merge": (("two-phase: all-gather of the ranks' [Q,k] score bounds, bound-limited rescore, " if not args.single_phase else "")
                          + "one packed NCCL all-gather of [dist|idx] + merge kernel") if world > 1 else "single shard",
                "n_settled_by_the_bound": stats.get("n_bound", 0),
                "n_certified": stats["n_certified"], "n_refused_first_pass": stats["n_fallback"],
                "n_resolved_by_stage2": stats["n_stage2"], "n_exact_scan": stats["n_scan"],
                "pipeline_chunks": stats["chunks"], "index_build_from_host_s": m["build_s"]})
    line = {
        "metric": "top-25 queries/s vs 1Mx4096 db", "value": Q / (ms * 1e-3), "unit": "queries/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f16 x f16 -> f32 (tcgen05 candidate pass) + f64 exact rescore", "data": "synthetic",
        "config": cfg,
        "roofline": {"bound": "tensor", "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / pk["tf_sustained"], "frac_of_burst_peak": achieved / pk["tf_burst"],
                     "peak_source": f"{pk['source']} cuBLAS bf16 sustained (kernel runs ~{tc_avg_ms:.0f} ms back to back under the power cap)",
                     "kernel": "knn_tc_kernel", "kernel_ms": tc_avg_ms, "kernel_share_of_step": tc_avg_ms / ms,
                     "whole_step_frac": m["flops_per_launch"] / (ms * 1e-3) / 1e12 / pk["tf_sustained"],
                     "algorithmic_flops_per_launch": m["flops_per_launch"],
                     # dram__bytes_read.sum + dram__bytes_write.sum over the launches of one step at the full 1M-row size,
                     # parsed from the committed `ncu --set full` digest (tensor-bound kernel: this is context -- the fp16
                     # shard is streamed ~4x per step -- not the roofline numerator)
                     "traffic": profile_traffic("r2_ncu_knn_tc.txt") if (m["rows_local"] == R_FULL and Q == Q_STEP) else None},
        "e2e": {"value": Q / (ms_e2e * 1e-3), "unit": "queries/s", "h2d_bytes_per_step": Q * D * 4 // world if world > 1 else Q * D * 4,
                "d2h_bytes_per_step": Q * k * 16, "ms_per_step": ms_e2e},
        "gpu_launches": m["launches_per_step"] * args.steps,
        "clocks": m["clocks"],
    }
    return line


def bench_config5(args, torch, dist_mod, rank, world, pk):
    """BASELINE config 5 (the north-star target): the database grows with the machine -- 1M x 4096 fp32 rows PER GPU, i.e.
    8M rows on 8 GPUs (16.4 GB fp32 + 8.2 GB fp16 shadow per GPU; one GPU cannot hold 8M x 4096 fp32 = 131 GB plus a
    65.5 GB shadow in 180 GB) -- and every query meets every row: queries replicated, per-shard exact top-25, one packed
    NCCL all-gather, merge.  A subset of the 100k queries is timed (10 000 per step; throughput is linear in Q)."""
    R5 = R_FULL * world
    m = retrieval_run(args, torch, dist_mod, rank, world, R5, Q_STEP, data="gaussian", e2e=True, clocks=False)
    ach = m["flops_per_launch"] / (m["tc_ms"] * 1e-3) / 1e12 if m["tc_ms"] > 0 else 0.0
    step = m["flops_per_launch"] / (m["ms"] * 1e-3) / 1e12
    return {"metric": f"top-25 queries/s vs {R5}x4096 db (config 5, db sharded over {world} GPUs)", "value": Q_STEP / (m["ms"] * 1e-3),
            "unit": "queries/s", "n_gpus": world, "ms_per_step": m["ms"], "scaling": "weak (rows per GPU fixed)",
            "config": dict(workload_config(R5, Q_STEP, world), rows_per_gpu=m["rows_local"], exactness=m["stats"],
                           query_subset=f"{Q_STEP} of config 5's 100k queries per step"),
            "roofline": {"bound": "tensor", "achieved_per_gpu": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": ach / pk["tf_sustained"], "whole_step_frac_per_gpu": step / pk["tf_sustained"],
                         "kernel": "knn_tc_kernel", "kernel_ms": m["tc_ms"], "kernel_share_of_step": m["tc_ms"] / m["ms"]},
            "e2e": {"value": Q_STEP / (m["ms_e2e"] * 1e-3), "unit": "queries/s", "ms_per_step": m["ms_e2e"]},
            "gpu_launches": m["launches_per_step"] * m["steps"]}


def bench_clustered(args, torch, pk, gaussian_ms):
    """Retrieval on clustered descriptors (consecutive frames + stops, synth.trajectory_*): how many queries the first
    tensor pass refuses, who resolves them (second tensor stage vs float64 scan) and what the step costs next to the
    Gaussian headline."""
    m = retrieval_run(args, torch, None, 0, 1, R_FULL, Q_STEP, data="clustered", e2e=False, clocks=False,
                      steps=max(3, min(args.steps, 5)), warmup=3)
    st = m["stats"]
    return {"metric": "top-25 queries/s vs 1Mx4096 db, clustered descriptors", "value": Q_STEP / (m["ms"] * 1e-3), "unit": "queries/s",
            "ms_per_step": m["ms"], "step_time_vs_gaussian": m["ms"] / gaussian_ms,
            "config": {"workload": "as the headline, db = 1M consecutive frames of one drive (great-circle path between random "
                                   "anchors every 64 frames, 6 % of the frames in stops of 100-300 near-identical frames), queries = "
                                   "perturbed frames", "exactness": st, **m["info"]},
            "roofline": {"bound": "tensor", "kernel_ms_first_pass": m["tc_ms"]},
            "gpu_launches": m["launches_per_step"] * m["steps"]}


def bench_wms(args, torch, pk, T=4096):
    from soft_contrastive_learning_b200 import losses, synth
    S, D = 25, D_FULL
    emb_s, dist_s, _ = synth.wms_batch(T=64, P=12, N=12, D=D, seed=42)        # 64 distinct tuples, tiled to T
    reps = T // 64
    emb = torch.tensor(emb_s, device="cuda").repeat(reps, 1, 1).contiguous()
    dist = torch.tensor(dist_s, device="cuda").repeat(reps, 1, 1).contiguous()
    emb += 1e-3 * torch.randn_like(emb)
    params = losses._ms_params(0.8, 15.0)

    def step():
        losses._wms_tuple_raw(emb, dist, params, need_grad=True)

    ms = timed(torch, step, max(args.steps, 10), max(args.warmup, 3))
    bytes_alg = T * (2 * S * D * 4 + S * S * 4)
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    # config 1 exactly (T=32): latency of one fused launch
    e32, d32 = emb[:32].contiguous(), dist[:32].contiguous()
    ms32 = timed(torch, lambda: losses._wms_tuple_raw(e32, d32, params, need_grad=True), 50, 10)
    # the same launch replayed from a CUDA graph (20 launches per graph): what a captured training step pays
    ms32_graph = None
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                losses._wms_tuple_raw(e32, d32, params, need_grad=True)
        torch.cuda.current_stream().wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for _ in range(20):
                keep = losses._wms_tuple_raw(e32, d32, params, need_grad=True)
        ms32_graph = timed(torch, graph.replay, 20, 5) / 20
        del graph, keep
    except Exception as e:                 # never cost the line
        print("cuda graph replay failed:", repr(e)[:200], file=sys.stderr)
    # end to end from pinned host buffers
    eh = torch.empty(emb.shape, dtype=torch.float32, pin_memory=True)
    eh.copy_(emb)
    dh = torch.empty(dist.shape, dtype=torch.float32, pin_memory=True)
    dh.copy_(dist)
    gh = torch.empty(emb.shape, dtype=torch.float32, pin_memory=True)
    lh = torch.empty(1, dtype=torch.float32, pin_memory=True)

    def step_e2e():
        e = eh.to("cuda", non_blocking=True)
        d = dh.to("cuda", non_blocking=True)
        loss, grad, _, _ = losses._wms_tuple_raw(e, d, params, need_grad=True)
        gh.copy_(grad, non_blocking=True)
        lh.copy_(loss, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    ms_e2e = timed(torch, step_e2e, max(3, args.steps), 3)
    cb = wms_cpu_port()
    return {"metric": "wms loss fwd+bwd tuples/s", "value": T / (ms * 1e-3), "unit": "tuples/s", "ms_per_step": ms,
            "dtype": "f32", "config": {"workload": f"wms tuple mode, T={T} S={S} D={D} fp32 (config 1 shape x{T // 32}; inputs 3.4 GB > L2)",
                                       "config1_T32_us_per_launch": ms32 * 1e3, "config1_T32_tuples_per_s": 32 / (ms32 * 1e-3),
                                       "config1_T32_us_per_launch_cuda_graph": None if ms32_graph is None else ms32_graph * 1e3,
                                       "config1_T32_hbm_frac_cuda_graph": None if ms32_graph is None else
                                       32 * (2 * S * D * 4 + S * S * 4) / (ms32_graph * 1e-3) / 1e9 / pk["hbm_gbs"]},
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": gbs / pk["hbm_gbs"],
                         "peak_source": pk["source"], "kernel": "wms_stream_kernel<5,7,224,packed Gram,tensor-core backward>",
                         "algorithmic_bytes_per_tuple": 2 * S * D * 4 + S * S * 4,
                         # dram read + write of one launch (T=4096), parsed from the committed `ncu --set full` digest
                         "traffic": profile_traffic("r2_ncu_wms.txt") if T == 4096 else None},
            "e2e": {"value": T / (ms_e2e * 1e-3), "unit": "tuples/s", "h2d_bytes_per_step": int(emb.numel() * 4 + dist.numel() * 4),
                    "d2h_bytes_per_step": int(emb.numel() * 4 + 4)},
            "cpu_baseline": cb, "gpu_launches": max(args.steps, 10)}


def _time_ms(torch, fn, iters, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_wms_sharded(args, torch, dist_mod, rank, world, pk, T_local=4096):
    """SURVEY 8e row 2 (N > 1): tuples sharded over the ranks, T_local per GPU (weak scaling), no data-path collective;
    each step ends with the scalar all-reduce of sharded.combine_tuple_shards.  value = tuples of ALL ranks / max time."""
    from soft_contrastive_learning_b200 import losses, sharded, synth
    S, D = 25, D_FULL
    emb_s, dist_s, _ = synth.wms_batch(T=64, P=12, N=12, D=D, seed=42 + rank)
    emb = torch.tensor(emb_s, device="cuda").repeat(T_local // 64, 1, 1).contiguous()
    dmat = torch.tensor(dist_s, device="cuda").repeat(T_local // 64, 1, 1).contiguous()
    emb += 1e-3 * torch.randn_like(emb)
    params = losses._ms_params(0.8, 15.0)

    def step():
        loss, grad, _, _ = losses._wms_tuple_raw(emb, dmat, params, need_grad=True)
        sharded.combine_tuple_shards(loss.reshape(()), None, T_local)       # gradients stay local: only the scalar travels

    ms = timed(torch, step, max(args.steps, 10), max(args.warmup, 3), dist_mod)
    bytes_alg = world * T_local * (2 * S * D * 4 + S * S * 4)
    gbs = bytes_alg / (ms * 1e-3) / 1e9
    return {"metric": "wms loss fwd+bwd tuples/s", "value": world * T_local / (ms * 1e-3), "unit": "tuples/s", "ms_per_step": ms,
            "n_gpus": world, "scaling": "weak", "dtype": "f32",
            "config": {"workload": f"wms tuple mode sharded by tuples, T={T_local} per GPU x {world} GPUs, S={S} D={D}; "
                                   "one scalar all-reduce per step, no data-path collective"},
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"] * world, "unit": "GB/s",
                         "frac": gbs / (pk["hbm_gbs"] * world), "peak_source": pk["source"] + f" x {world} GPUs"},
            "gpu_launches": max(args.steps, 10)}


def bench_netvlad_sharded(args, torch, dist_mod, rank, world, pk, B_local=32, H=30, W=40, Cc=512, K=64):
    """SURVEY 8e row 4 (N > 1): the images of the batch split over the ranks (B_local per GPU, weak scaling), weights
    replicated, forward + backward per rank and ONE NCCL all-reduce of the packed [dW | dC]."""
    from soft_contrastive_learning_b200 import sharded
    g = torch.Generator(device="cuda").manual_seed(42 + rank)
    x = torch.randn((B_local, H * W, Cc), generator=g, device="cuda")
    gw = torch.Generator(device="cuda").manual_seed(1)
    aw = 0.05 * torch.randn((Cc, K), generator=gw, device="cuda")
    cc = 0.05 * torch.randn((Cc, K), generator=gw, device="cuda")
    dout = torch.randn((B_local, Cc * K), generator=g, device="cuda") / (B_local * world)
    out = {}

    def step():
        out["r"] = sharded.netvlad_step_sharded(x, aw, cc, lambda v: dout)

    ms = timed(torch, step, max(args.steps, 10), 3, dist_mod)
    dw = out["r"][2]
    chk = dw.clone()
    dist_mod.all_reduce(chk, op=dist_mod.ReduceOp.MAX)
    same = bool(torch.equal(chk, dw))
    return {"metric": "NetVLAD head fwd+bwd images/s, batch-parallel", "value": world * B_local / (ms * 1e-3), "unit": "images/s",
            "ms_per_step": ms, "n_gpus": world, "scaling": "weak",
            "config": {"workload": f"B={B_local} images per GPU x {world} GPUs, {H}x{W}x{Cc} maps, K={K}; one all-reduce of [dW|dC] "
                                   f"({2 * Cc * K * 4} bytes) per step", "grads_identical_on_all_ranks": same},
            "gpu_launches": max(args.steps, 10)}


def bench_netvlad_pca(args, torch, pk, B=256, H=30, W=40, Cc=512, K=64, Dout=4096):
    """BASELINE config 2: NetVLAD head K=64 over 30x40x512 maps + PCA 32768->4096, forward and backward, batch 256."""
    import ctypes as C
    from soft_contrastive_learning_b200._lib import check, lib
    from soft_contrastive_learning_b200.losses import _p, _stream, _ws
    L = lib()
    HW, Din = H * W, Cc * K
    g = torch.Generator(device="cuda").manual_seed(42)
    x = torch.randn((B, HW, Cc), generator=g, device="cuda")
    aw = 0.05 * torch.randn((Cc, K), generator=g, device="cuda")
    cc = 0.05 * torch.randn((Cc, K), generator=g, device="cuda")
    V = torch.randn((Dout, Din), generator=g, device="cuda") / Din ** 0.5
    m = 0.01 * torch.randn(Din, generator=g, device="cuda")
    var = 0.5 + 1.5 * torch.rand(Dout, generator=g, device="cuda")
    n = C.c_size_t()
    check(L.scl_netvlad_workspace_bytes(B, HW, Cc, K, C.byref(n)), "ws")
    ws = _ws(n.value, x.device)
    check(L.scl_pca_workspace_bytes(B, Din, Dout, C.byref(n)), "ws")
    pws = _ws(n.value, x.device)
    vlad = torch.empty((B, Din), device="cuda")
    y = torch.empty((B, Dout), device="cuda")
    dy = torch.randn((B, Dout), generator=g, device="cuda")
    dvlad = torch.empty((B, Din), device="cuda")
    dx, dw, dc = torch.empty_like(x), torch.empty_like(aw), torch.empty_like(cc)
    st = _stream()
    f_nv = lambda: check(L.scl_netvlad_fwd(_p(x), _p(aw), _p(cc), B, HW, Cc, K, _p(vlad), _p(ws), ws.numel(), st), "nv fwd")
    f_pf = lambda: check(L.scl_pca_fwd(_p(vlad), _p(V), _p(m), _p(var), B, Din, Dout, _p(y), _p(pws), pws.numel(), st), "pca fwd")
    f_pb = lambda: check(L.scl_pca_bwd(_p(dy), _p(V), _p(var), B, Din, Dout, _p(dvlad), _p(pws), pws.numel(), st), "pca bwd")
    f_nb = lambda: check(L.scl_netvlad_bwd(_p(x), _p(aw), _p(cc), _p(dvlad), B, HW, Cc, K, _p(dx), _p(dw), _p(dc), _p(ws),
                                           ws.numel(), st), "nv bwd")
    # the PCA matrix is a fed constant (train/train.py:281-283): split once into fp16 hi / lo halves, outside the step
    check(L.scl_pca_shadow_bytes(Din, Dout, C.byref(n)), "shadow bytes")
    shadow = _ws(n.value, x.device)
    f_prep = lambda: check(L.scl_pca_prepare(_p(V), Din, Dout, _p(shadow), shadow.numel(), st), "pca prepare")
    prep_ms = _time_ms(torch, f_prep, 3, 1)
    check(L.scl_pca_prepared_workspace_bytes(B, Din, Dout, C.byref(n)), "ws")
    qws = _ws(n.value, x.device)
    f_pfp = lambda: check(L.scl_pca_fwd_prepared(_p(vlad), _p(shadow), _p(m), _p(var), B, Din, Dout, _p(y), _p(qws), qws.numel(), st), "pca fwd prepared")
    f_pbp = lambda: check(L.scl_pca_bwd_prepared(_p(dy), _p(shadow), _p(var), B, Din, Dout, _p(dvlad), _p(qws), qws.numel(), st), "pca bwd prepared")
    out = {}
    for prec, tag in ((0, "fp32-grade 3xTF32"), (1, "single-pass TF32")):
        check(L.scl_set_gemm_precision(prec), "prec")
        t = {"netvlad_fwd": _time_ms(torch, f_nv, 5), "pca_fwd": _time_ms(torch, f_pf, 5), "pca_bwd": _time_ms(torch, f_pb, 5),
             "netvlad_bwd": _time_ms(torch, f_nb, 5)}
        if prec == 0:
            # fp32-grade mode as the host wrapper runs it: fused NetVLAD forward (fp16 hi/lo operands in tensor memory) and
            # the prepared PCA (pre-split f16 engine); the in-kernel 3xTF32 split is kept beside it
            t["pca_fwd_3xtf32_unprepared"], t["pca_bwd_3xtf32_unprepared"] = t["pca_fwd"], t["pca_bwd"]
            t["pca_fwd"], t["pca_bwd"] = _time_ms(torch, f_pfp, 5), _time_ms(torch, f_pbp, 5)
            t["pca_prepare_once"] = prep_ms
        t["total"] = t["netvlad_fwd"] + t["pca_fwd"] + t["pca_bwd"] + t["netvlad_bwd"]
        out[tag] = t
    check(L.scl_set_gemm_precision(0), "prec")
    flops = {"netvlad_fwd": 2 * 2.0 * B * HW * Cc * K, "netvlad_bwd": 4 * 2.0 * B * HW * Cc * K,
             "pca_fwd": 2.0 * B * Din * Dout, "pca_bwd": 2.0 * B * Din * Dout}
    t0 = out["fp32-grade 3xTF32"]
    tot_flops = sum(flops.values())
    ach = tot_flops / (t0["total"] * 1e-3) / 1e12
    # algorithmic HBM bytes of the step: X read by the forward and by the backward, dX written, V read by
    # both PCA passes, the [B,32768] VLAD / gradient and the [B,4096] outputs once each
    hbm_bytes = 4.0 * (3 * B * HW * Cc + 2 * Dout * Din + 4 * B * Din + 2 * B * Dout)
    hbm_gbs = hbm_bytes / (t0["total"] * 1e-3) / 1e9
    return {"metric": "NetVLAD head + PCA fwd+bwd images/s", "value": B / (t0["total"] * 1e-3), "unit": "images/s",
            "ms_per_step": t0["total"], "dtype": "fp32-grade on tcgen05: fp16 hi/lo split x3 (NetVLAD forward, PCA) and tf32 x3 (NetVLAD backward), fp32 accumulate",
            "config": {"workload": f"BASELINE config 2: B={B}, {H}x{W}x{Cc} conv5 maps, K={K}, PCA {Din}->{Dout}, fwd+bwd",
                       "ms": out, "inputs": "x 629 MB + V 537 MB fp32 (>> L2)"},
            "roofline": {"bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": ach / pk["tf_sustained"], "peak_source": pk["source"] + " cuBLAS bf16 sustained (no tf32 peak measured; "
                         "kind::tf32 is nominally half the bf16 rate and the fp32-grade mode issues 3 MMAs per product)",
                         "algorithmic_flops_per_step": tot_flops,
                         "kernel": "nv_fused_kernel<0/1>, nv_dx_kernel, tc_gemm_h3_kernel", "traffic": None,
                         "hbm_view": {"algorithmic_bytes_per_step": hbm_bytes, "achieved_gbs": hbm_gbs, "peak_gbs": pk["hbm_gbs"],
                                      "frac": hbm_gbs / pk["hbm_gbs"],
                                      "note": "with fp32 X the step is HBM-bound, not tensor-bound (SURVEY 8d): this is the binding roofline"}},
            "gpu_launches": 5 * (6 + 11 + 4 + 2)}


def bench_losses_cfg3(args, torch, pk, T=32, P=15, N=16, D=D_FULL):
    """BASELINE config 3: every hot-path loss, forward+backward, on a 1024-descriptor batch (32 tuples x 32, D=4096)."""
    from soft_contrastive_learning_b200 import losses, synth
    rng = np.random.default_rng(42)
    S = 1 + P + N
    xy = synth.tuple_xy(rng, T, P, N)
    emb = synth.tuple_descriptors(rng, T, P, N, D)
    e3 = torch.tensor(emb, device="cuda")
    e2 = e3.reshape(T * S, D)
    d3 = torch.tensor(synth.pairwise_euclid(xy).astype(np.float32), device="cuda")
    xyf = xy.reshape(-1, 2)
    dflat = torch.tensor(np.sqrt(((xyf[:, None] - xyf[None]) ** 2).sum(-1)).astype(np.float32), device="cuda")
    sqd = torch.tensor(synth.anchor_sq_dists(xy, P).astype(np.float32), device="cuda")
    labels = losses.ms_labels(T, P, N)
    # logratio needs P == N (model/losses.py:125-135): 15 + 15 on the first 31 rows of each tuple
    e_lr = e3[:, :31].contiguous().reshape(T * 31, D)
    sp, sn = synth.logratio_sq_dists(xy[:, :31], 15, 15)
    sp, sn = torch.tensor(sp.astype(np.float32), device="cuda"), torch.tensor(sn.astype(np.float32), device="cuda")
    p = losses._ms_params(0.8, 15.0)
    runs = {
        "triplet": lambda: losses.tuple_loss_value_and_grad("triplet_loss", e2, T, P, N),
        "lazy_triplet": lambda: losses.tuple_loss_value_and_grad("lazy_triplet_loss", e2, T, P, N),
        "quadruplet": lambda: losses.tuple_loss_value_and_grad("quadruplet_loss", e2, T, P, N - 1),
        "lazy_quadruplet": lambda: losses.tuple_loss_value_and_grad("lazy_quadruplet_loss", e2, T, P, N - 1),
        "huber_distance_triplet": lambda: losses.tuple_loss_value_and_grad("triplet_loss", e2, T, P, N, squared_d_dists=sqd,
                                                                             distance_loss_name="huber_distance_loss"),
        "logratio": lambda: losses.logratio_loss_value_and_grad(e_lr, T, 15, 15, sp, sn),
        "ms_loss (flat, B=1024)": lambda: losses._flat_raw(e2, None, losses._labels_i32(labels, e2.device), losses._ms_params()),
        "wms (flat, B=1024)": lambda: losses._flat_raw(e2, dflat, None, p),
        "wms (tuples, T=32 S=32)": lambda: losses._wms_tuple_raw(e3, d3, p),
    }
    ms = {k: _time_ms(torch, f, 20, 5) for k, f in runs.items()}
    total = sum(ms.values())
    bytes_alg = 2.0 * T * S * D * 4
    return {"metric": "all hot-path losses fwd+bwd on a 1024-descriptor batch, sweeps/s", "value": 1e3 / total, "unit": "sweeps/s",
            "ms_per_step": total, "dtype": "f32",
            "config": {"workload": f"BASELINE config 3: T={T} tuples x S={S} (P={P}, N={N}; quadruplet N={N - 1}+1; logratio 15+15), D={D}",
                       "ms_per_loss": ms, "note": "33.5 MB in+out per loss: launch-latency bound at this size (4 us at HBM peak)"},
            "roofline": {"bound": "hbm", "achieved": len(ms) * bytes_alg / (total * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": len(ms) * bytes_alg / (total * 1e-3) / 1e9 / pk["hbm_gbs"], "peak_source": pk["source"],
                         "algorithmic_bytes_per_loss": bytes_alg, "traffic": None},
            "gpu_launches": 20 * len(ms)}


def main():
    # libraries (NCCL's version banner) may write to fd 1: keep the real stdout for the ONE JSON line, send the rest to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="retrieval", choices=["retrieval", "wms", "netvlad", "losses"])
    ap.add_argument("--rows", type=int, default=R_FULL)
    ap.add_argument("--queries", type=int, default=Q_STEP)
    ap.add_argument("--data", default="gaussian", choices=["gaussian", "clustered"])
    ap.add_argument("--single-phase", action="store_true",
                    help="N > 1: the plain sharded protocol (every rank returns its full local top-k) instead of the two-phase one")
    ap.add_argument("--group-pipeline", action="store_true",
                    help="N > 1, two-phase: second phase per query group on a second stream under the tensor launch")
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--time-build", action="store_true", default=True)
    args = ap.parse_args()
    args.warmup = max(3, args.warmup)

    if args.impl == "reference":
        run_reference(args, real_stdout)
        return

    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist_mod = None
    if world > 1:
        import torch.distributed as dist_mod
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
    pk = peaks()
    if args.workload in ("netvlad", "losses"):
        line = (bench_netvlad_pca if args.workload == "netvlad" else bench_losses_cfg3)(args, torch, pk)
        line.update({"n_gpus": 1, "steps": 5, "warmup": 3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                     "data": "synthetic"})
    elif args.workload == "wms":
        line = bench_wms(args, torch, pk)
        line.update({"n_gpus": 1, "steps": max(args.steps, 10), "warmup": max(args.warmup, 3), "higher_is_better": True,
                     "scaling": "weak", "vs_baseline": None, "data": "synthetic"})
    else:
        line = bench_retrieval(args, torch, dist_mod, rank, world, pk)
        cfg = line["config"]          # flat scalar keys: the driver's parser keeps them (it drops `secondary`)
        if rank == 0 and world == 1:
            if not args.no_cpu_baseline:
                line["cpu_baseline"] = kdtree_reference(steps=1, warmup=0)
                line["cpu_baseline"].pop("ms_per_step", None)
            if not args.no_secondary:
                line["secondary"] = []

                def secondary(fn, *a):
                    torch.cuda.empty_cache()
                    try:
                        r = fn(*a)
                    except Exception as e:          # a secondary line must never cost the headline
                        r = {"metric": fn.__name__, "error": repr(e)[:300]}
                    line["secondary"].append(r)
                    return r
                w = secondary(bench_wms, args, torch, pk)
                if "error" not in w:
                    # metric B of BASELINE.json as a top-level object (the same numbers also sit in `config` as flat keys)
                    line["wms"] = {"t4096": {"tuples_per_s": w["value"], "ms_per_step": w["ms_per_step"], "roofline": w["roofline"],
                                             "e2e": w["e2e"], "cpu_baseline": w["cpu_baseline"]},
                                   "config1_t32": {"us_per_launch": w["config"]["config1_T32_us_per_launch"],
                                                   "us_per_launch_cuda_graph": w["config"]["config1_T32_us_per_launch_cuda_graph"],
                                                   "tuples_per_s": w["config"]["config1_T32_tuples_per_s"],
                                                   "hbm_frac_cuda_graph": w["config"]["config1_T32_hbm_frac_cuda_graph"]}}
                    cfg.update({"wms_t4096_tuples_per_s": w["value"], "wms_t4096_ms": w["ms_per_step"],
                                "wms_t4096_hbm_frac": w["roofline"]["frac"],
                                "wms_config1_t32_us_per_launch": w["config"]["config1_T32_us_per_launch"],
                                "wms_config1_t32_us_cuda_graph": w["config"]["config1_T32_us_per_launch_cuda_graph"],
                                "wms_e2e_tuples_per_s": w["e2e"]["value"], "wms_cpu_port_tuples_per_s": w["cpu_baseline"]["value"]})
                nv = secondary(bench_netvlad_pca, args, torch, pk)
                if "error" not in nv:
                    t = nv["config"]["ms"]["fp32-grade 3xTF32"]
                    cfg.update({"cfg2_netvlad_pca_ms_per_step": nv["ms_per_step"], "cfg2_netvlad_fwd_ms": t["netvlad_fwd"],
                                "cfg2_netvlad_bwd_ms": t["netvlad_bwd"], "cfg2_pca_fwd_ms": t["pca_fwd"], "cfg2_pca_bwd_ms": t["pca_bwd"],
                                "cfg2_hbm_frac": nv["roofline"]["hbm_view"]["frac"], "cfg2_images_per_s": nv["value"]})
                ls = secondary(bench_losses_cfg3, args, torch, pk)
                if "error" not in ls:
                    cfg.update({"cfg3_all_losses_sweep_ms": ls["ms_per_step"],
                                "cfg3_triplet_us": ls["config"]["ms_per_loss"]["triplet"] * 1e3,
                                "cfg3_wms_flat_b1024_us": ls["config"]["ms_per_loss"]["wms (flat, B=1024)"] * 1e3})
                if args.data == "gaussian" and args.rows == R_FULL:
                    cl = secondary(bench_clustered, args, torch, pk, line["ms_per_step"])
                    if "error" not in cl:
                        st = cl["config"]["exactness"]
                        cfg.update({"clustered_queries_per_s": cl["value"], "clustered_ms_per_step": cl["ms_per_step"],
                                    "clustered_step_time_vs_gaussian": cl["step_time_vs_gaussian"],
                                    "clustered_n_refused_first_pass": st["n_fallback"], "clustered_n_resolved_by_stage2": st["n_stage2"],
                                    "clustered_n_exact_scan": st["n_scan"]})
        if world > 1 and not args.no_secondary:
            # a secondary line must never cost the headline: if a rank fails and the others wait in a collective, a
            # watchdog prints the headline alone and ends the process
            line["secondary"] = []

            def bail():
                if rank == 0:
                    line["secondary"].append({"metric": "secondary", "error": "timed out"})
                    print(json.dumps(line), file=real_stdout, flush=True)
                os._exit(0)
            dog = threading.Timer(300.0, bail)
            dog.daemon = True
            dog.start()
            torch.cuda.empty_cache()
            try:
                c5 = bench_config5(args, torch, dist_mod, rank, world, pk)
                cfg.update({"cfg5_rows": R_FULL * world, "cfg5_queries_per_s": c5["value"], "cfg5_ms_per_step": c5["ms_per_step"],
                            "cfg5_tensor_frac_per_gpu": c5["roofline"]["frac"],
                            "cfg5_whole_step_frac_per_gpu": c5["roofline"]["whole_step_frac_per_gpu"],
                            "cfg5_kernel_share_of_step": c5["roofline"]["kernel_share_of_step"],
                            "cfg5_e2e_queries_per_s": c5["e2e"]["value"], "cfg5_n_refused": c5["config"]["exactness"]["n_fallback"]})
            except Exception as e:
                c5 = {"metric": "bench_config5", "error": repr(e)[:300]}
            line["secondary"].append(c5)
            torch.cuda.empty_cache()
            try:
                sec = bench_wms_sharded(args, torch, dist_mod, rank, world, pk)
                cfg.update({"wms_sharded_tuples_per_s": sec["value"], "wms_sharded_hbm_frac": sec["roofline"]["frac"]})
            except Exception as e:
                sec = {"metric": "bench_wms_sharded", "error": repr(e)[:300]}
            line["secondary"].append(sec)
            try:
                nvs = bench_netvlad_sharded(args, torch, dist_mod, rank, world, pk)
                cfg.update({"netvlad_dp_images_per_s": nvs["value"], "netvlad_dp_ms_per_step": nvs["ms_per_step"]})
            except Exception as e:
                nvs = {"metric": "bench_netvlad_sharded", "error": repr(e)[:300]}
            line["secondary"].append(nvs)
            dog.cancel()
    if rank == 0:
        print(json.dumps(line), file=real_stdout, flush=True)
    if dist_mod is not None:
        dist_mod.barrier()
        dist_mod.destroy_process_group()


if __name__ == "__main__":
    main()
