set -x
timeout 900 python -m pytest tests/test_gpu_retrieval.py -x -q -m gpu 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:knn_cand_merge -s 2 -c 1 -o gpurun_out/r2_cand_merge python tools/knn_merge_prof.py > gpurun_out/cand_merge_ncu.log 2>&1
tail -3 gpurun_out/cand_merge_ncu.log
