#!/bin/bash
# DRAM traffic / L2 hit rate / tensor-pipe activity of the tensor pass for a list of env configurations.
mkdir -p gpurun_out
: > gpurun_out/l2_sweep.txt
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,l1tex__m_xbar2l1tex_read_bytes.sum
for cfg in "$@"; do
  env $cfg ncu --metrics $M --clock-control none -k regex:knn_tc_kernel -s 2 -c 1 --csv --log-file gpurun_out/l2_one.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2> gpurun_out/l2_one.err
  echo "== $cfg" >> gpurun_out/l2_sweep.txt
  python - <<'PY' >> gpurun_out/l2_sweep.txt
import csv
rows = [r for r in csv.reader(open("gpurun_out/l2_one.csv")) if len(r) > 5]
hdr = rows[0]
ix = {h: i for i, h in enumerate(hdr)}
for r in rows[1:]:
    print("  ", r[ix["Metric Name"]], r[ix["Metric Value"]], r[ix["Metric Unit"]])
PY
done
cat gpurun_out/l2_sweep.txt
