#!/bin/bash
# Round evidence: (1) launch list of the default bench, (2) full-set captures of the dominant kernels, (3) clocks.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:knn_tc_kernel -s 3 -c 1 -f -o gpurun_out/prof_knn_tc \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/ncu_tc.log 2>&1
echo "tc capture rc=$?"
ncu --set full --clock-control none --import-source on -k regex:wms_stream_kernel -s 5 -c 1 -f -o gpurun_out/prof_wms \
    python bench.py --workload wms --steps 3 --warmup 3 > gpurun_out/ncu_wms.log 2>&1
echo "wms capture rc=$?"
# config 2: the fused NetVLAD forward / backward kernels, the dx kernel and the pre-split f16 GEMM of the prepared PCA
for k in "nv_fused_kernel<0>:nvfwd" "nv_fused_kernel<1>:nvbwd" "nv_dx_kernel:nvdx" "tc_gemm_h3_kernel:h3"; do
  ncu --set full --clock-control none --import-source on -k "regex:${k%%:*}" -s 3 -c 1 -f -o gpurun_out/prof_${k##*:} \
      python bench.py --workload netvlad > gpurun_out/ncu_${k##*:}.log 2>&1
  echo "${k##*:} capture rc=$?"
done
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap \
    --format=csv -lms 100 > gpurun_out/clocks.csv &
SMI=$!
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
kill $SMI
ls -la gpurun_out/*.ncu-rep
