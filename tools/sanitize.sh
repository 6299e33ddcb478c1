#!/bin/bash
# compute-sanitizer over one small shape per kernel family.  Run on the GPU box:
#   gpurun --timeout 1500 -- 'bash tools/sanitize.sh'
# Logs land in gpurun_out/sanitize_{memcheck,racecheck}_<family>.log; copy the summaries to profiles/.
mkdir -p gpurun_out
for tool in memcheck racecheck; do
  for fam in wms tuples flat netvlad knn; do
    timeout 600 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py $fam \
      > gpurun_out/sanitize_${tool}_${fam}.log 2>&1
    echo "$tool $fam rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/sanitize_${tool}_${fam}.log | tail -1)"
  done
done | tee gpurun_out/sanitize_summary.txt
