#!/bin/bash
# ncu evidence: (1) launch list of a short bench run, (2) full-set capture of the dominant kernels, (3) clocks under load.
mkdir -p gpurun_out
ROWS=${ROWS:-1000000}
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 3 --rows $ROWS --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
for v in ${VARIANTS:-3}; do
SCL_KNN_TC_VARIANT=$v ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_knn_tc_v$v \
    python bench.py --steps 1 --warmup 3 --rows $ROWS --no-cpu-baseline --no-secondary > gpurun_out/ncu_tc_v$v.log 2>&1
echo "tc capture v$v rc=$?"
done
if [ -z "$SKIP_WMS" ]; then
ncu --set full --clock-control none --import-source on -k regex:wms_tuple -s 5 -c 1 -f -o gpurun_out/prof_wms \
    python bench.py --workload wms --steps 3 --warmup 3 > gpurun_out/ncu_wms.log 2>&1
echo "wms capture rc=$?"
fi
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 100 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --steps 20 --warmup 3 --rows $ROWS --no-cpu-baseline --no-secondary > gpurun_out/bench_clocks.json 2> gpurun_out/bench_clocks.err
kill $SMI
ls -la gpurun_out/*.ncu-rep
