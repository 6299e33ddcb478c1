#!/bin/bash
# tensor-pass time at the shard sizes of N = 8, 4, 1 on one GPU
for r in ${ROWS_LIST:-125000 250000 1000000}; do
  timeout 200 python bench.py --rows $r --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/b5_$r.json 2> gpurun_out/b5.err
  python - $r <<'PY'
import json, sys
s = open("gpurun_out/b5_" + sys.argv[1] + ".json").read()
j = json.loads(s[s.index("{"):])
print(sys.argv[1], round(j["value"]), "q/s ms", round(j["ms_per_step"], 2), "kernel", round(j["roofline"]["kernel_ms"], 2), round(j["roofline"]["frac"], 3), j["config"]["exactness"])
PY
done
