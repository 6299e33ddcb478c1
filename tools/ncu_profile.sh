#!/bin/bash
# ncu evidence: (1) launch list of a short bench run, (2) full-set capture of the two dominant kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --rows ${ROWS:-200000} --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_knn_tc \
    python bench.py --steps 1 --warmup 3 --rows ${ROWS:-200000} --no-cpu-baseline --no-secondary > gpurun_out/ncu_tc.log 2>&1
echo "tc capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:wms_tuple_kernel -s 5 -c 1 -f -o gpurun_out/prof_wms \
    python bench.py --workload wms --steps 3 --warmup 3 > gpurun_out/ncu_wms.log 2>&1
echo "wms capture rc=$?"
ls -la gpurun_out/*.ncu-rep
