#!/bin/bash
# Sweep tensor-pass variants / group sizes / range counts on the full 1M x 4096 problem.
mkdir -p gpurun_out
: > gpurun_out/sweep.txt
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/sw.json 2> gpurun_out/sw.err
  python - "$cfg" <<'PY' >> gpurun_out/sweep.txt
import json, sys
try:
    j = json.load(open("gpurun_out/sw.json"))
    print(sys.argv[1], "| q/s", round(j["value"]), "ms", round(j["ms_per_step"], 2), "TF", round(j["roofline"]["achieved"], 1),
          "frac", round(j["roofline"]["frac"], 3), "kernel_ms", round(j["roofline"]["kernel_ms"], 2), "e2e", round(j["e2e"]["value"]),
          "clk", j["clocks"]["sm_mhz"], j["clocks"]["reasons"], "fb", j["config"]["exactness"]["n_fallback"])
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/sw.err").read()[-500:])
PY
done
cat gpurun_out/sweep.txt
