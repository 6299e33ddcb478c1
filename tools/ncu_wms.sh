#!/bin/bash
# DRAM traffic / duration of the wms streaming kernel for a list of env configurations.
mkdir -p gpurun_out
: > gpurun_out/wms_sweep.txt
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
for cfg in "$@"; do
  env $cfg ncu --metrics $M --clock-control none -k regex:wms_stream_kernel -s 5 -c 1 --csv --log-file gpurun_out/wms_one.csv \
      python bench.py --workload wms --steps 3 --warmup 3 > /dev/null 2> gpurun_out/wms_one.err
  env $cfg python bench.py --workload wms --steps 10 --warmup 3 > gpurun_out/wms_one.json 2>> gpurun_out/wms_one.err
  echo "== $cfg" >> gpurun_out/wms_sweep.txt
  python - <<'PY' >> gpurun_out/wms_sweep.txt
import csv, json
rows = [r for r in csv.reader(open("gpurun_out/wms_one.csv")) if len(r) > 5]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
for r in rows[1:]:
    print("  ", r[ix["Metric Name"]], r[ix["Metric Value"]], r[ix["Metric Unit"]])
try:
    j = json.load(open("gpurun_out/wms_one.json"))
    print("   bench: ms", round(j["ms_per_step"], 4), "frac", round(j["roofline"]["frac"], 3), "tuples/s", round(j["value"]))
except Exception as e:
    print("   bench failed", e)
PY
done
cat gpurun_out/wms_sweep.txt
