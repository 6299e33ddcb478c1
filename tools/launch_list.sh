#!/bin/bash
# per-kernel device time of one retrieval step at a given shard size
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r.csv python bench.py --rows ${ROWS:-125000} --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/launches_r.csv")))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, start = r, i
        break
ix = {h: i for i, h in enumerate(hdr)}
seq = []
for r in rows[start + 2:]:
    if len(r) < len(hdr):
        continue
    v = float(r[ix["Metric Value"]]); u = r[ix["Metric Unit"]]
    v = v / 1e3 if u == "us" else v / 1e6 if u == "ns" else v * 1e3 if u == "s" else v
    seq.append((r[ix["Kernel Name"]][:70], v))
idx = [i for i, (n, v) in enumerate(seq) if "knn_query_prep" in n]
for n, v in seq[idx[-2]:idx[-1]]:
    print(f"{v:9.4f} ms  {n}")
PY
