#!/usr/bin/env python
"""Regenerate the committed evidence under profiles/ from what tools/ncu_profile.sh left in gpurun_out/.

    python tools/make_profiles.py [round_tag]        # default r1

launch list  -> profiles/<tag>_launch_list_summary.txt   (per-kernel totals and shares of the default bench step)
clocks.csv   -> profiles/<tag>_clocks_summary.txt
*.ncu-rep    -> profiles/<tag>_ncu_{knn_tc,wms,netvlad_fused_fwd,netvlad_fused_bwd,netvlad_dx,gemm_h3_pca}.txt  (tools/ncu_digest.py)
bench line   -> profiles/<tag>_bench_full_n1.json
"""
import collections, csv, os, shutil, statistics, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")


def launch_list(tag):
    path = os.path.join(OUT, "launches.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr = rows[start]
    ix = {h: i for i, h in enumerate(hdr)}
    tot, cnt = collections.Counter(), collections.Counter()
    seq = []
    for r in rows[start + 1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v, u = float(r[ix["Metric Value"]]), r[ix["Metric Unit"]]
        ms = v / 1e3 if u in ("us", "usecond") else v / 1e6 if u in ("ns", "nsecond") else v * 1e3 if u in ("s", "second") else v
        tot[r[ix["Kernel Name"]]] += ms
        cnt[r[ix["Kernel Name"]]] += 1
        seq.append((r[ix["Kernel Name"]], ms))
    total = sum(tot.values())
    with open(os.path.join(PROF, f"{tag}_launch_list_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 2 --warmup 3 --no-cpu-baseline\n")
        f.write(f"(cold-cache, serialised launches: compare SHARES, not absolutes).  total device time {total:.1f} ms over "
                f"{sum(cnt.values())} launches\n\n  total ms  count     avg ms   share  kernel\n")
        for k, v in tot.most_common(25):
            f.write(f"{v:10.3f} {cnt[k]:6d} {v / cnt[k]:10.4f} {100 * v / total:6.1f}%  {k[:110]}\n")
        # one retrieval step in isolation: the launches between two consecutive query-prep kernels
        prep = [i for i, (n, _) in enumerate(seq) if "knn_query_prep" in n]
        if len(prep) >= 2:
            step = seq[prep[-2]:prep[-1]]
            st = sum(ms for _, ms in step)
            f.write(f"\none retrieval step (10 000 queries vs the 1M-row shard), {len(step)} launches, {st:.3f} ms device time:\n")
            for n, ms in step:
                f.write(f"{ms:10.4f} ms {100 * ms / st:6.1f}%  {n[:110]}\n")


def clocks(tag):
    path = os.path.join(OUT, "clocks.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    hdr = [h.strip() for h in rows[0]]
    ix = {h.split(" ")[0]: i for i, h in enumerate(hdr)}
    sm, pw, reasons, n = [], [], set(), 0
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        n += 1
        try:
            p = float(r[ix["power.draw"]].strip().split()[0])
            c = float(r[ix["clocks.current.sm"]].strip().split()[0])
        except ValueError:
            continue
        pw.append(p)
        if p > 600:
            sm.append(c)
        for key in ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"):
            col = next((i for h, i in ix.items() if h.endswith(key)), None)
            if col is not None and r[col].strip().lower() == "active":
                reasons.add(key)
    cmax = rows[1][ix["clocks.max.sm"]].strip()
    with open(os.path.join(PROF, f"{tag}_clocks_summary.txt"), "w") as f:
        f.write(f"nvidia-smi -lms 100 during `python bench.py --steps 10 --warmup 3` (full default run): {n} samples, "
                f"{len(sm)} under load (>600 W)\n")
        if sm:
            f.write(f"sm clock under load: median {statistics.median(sm)} MHz (min {min(sm)}, max {max(sm)}), clocks.max.sm {cmax}\n")
        f.write(f"power: max {max(pw):.0f} W; throttle reasons seen: {sorted(reasons)}\n")


def digests(tag):
    for rep, name in (("prof_knn_tc", "ncu_knn_tc"), ("prof_wms", "ncu_wms"), ("prof_gemm", "ncu_gemm"),
                      ("prof_nvfwd", "ncu_netvlad_fused_fwd"), ("prof_nvbwd", "ncu_netvlad_fused_bwd"),
                      ("prof_nvdx", "ncu_netvlad_dx"), ("prof_h3", "ncu_gemm_h3_pca")):
        path = os.path.join(OUT, rep + ".ncu-rep")
        if os.path.exists(path):
            out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_digest.py"), path], capture_output=True, text=True).stdout
            open(os.path.join(PROF, f"{tag}_{name}.txt"), "w").write(out)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    os.makedirs(PROF, exist_ok=True)
    launch_list(tag)
    clocks(tag)
    digests(tag)
    b = os.path.join(OUT, "bench_full.json")
    if os.path.exists(b):
        shutil.copy(b, os.path.join(PROF, f"{tag}_bench_full_n1.json"))


if __name__ == "__main__":
    main()
