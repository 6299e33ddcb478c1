#!/bin/bash
# Run each GPU test file in its own process with its own timeout so a hung kernel cannot take the others down.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
free -g >> gpurun_out/gpu.txt; nproc >> gpurun_out/gpu.txt
for f in test_gpu_losses test_gpu_netvlad test_gpu_retrieval test_gpu_threads; do
  timeout ${TEST_TIMEOUT:-900} python -m pytest tests/$f.py -q -m gpu -x --no-header -p no:cacheprovider ${PYTEST_EXTRA} > gpurun_out/$f.log 2>&1
  echo "$f exit=$?" | tee -a gpurun_out/summary.txt
  tail -n 30 gpurun_out/$f.log
done
