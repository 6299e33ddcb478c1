timeout 900 python -m pytest tests/test_gpu_retrieval.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python tools/knn_two_phase.py 8 2>&1 | grep -E "rank 0 step|time-line" -A 0 | head -4
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"knn_query_prep" -c 4 --csv python tools/knn_merge_prof.py 2>/dev/null | grep knn_query_prep | awk -F'","' '{print $5, $(NF)}'
