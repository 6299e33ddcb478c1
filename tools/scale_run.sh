#!/bin/bash
# 1 -> N scaling of the default bench on one box (what the driver does at round end).
mkdir -p gpurun_out
for n in ${NS:-1 2 4 8}; do
  if [ "$n" = "1" ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline --no-secondary > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  python - $n <<'PY'
import json, sys
n = sys.argv[1]
try:
    s = open(f"gpurun_out/scale_n{n}.json").read()
    j = json.loads(s[s.index("{"):])
    print(f"N={n}: {j['value']:.0f} q/s, {j['ms_per_step']:.2f} ms/step, kernel {j['roofline']['kernel_ms']:.2f} ms ({j['roofline']['frac']:.3f}), e2e {j['e2e']['value']:.0f}, clocks {j['clocks']['sm_mhz']} {j['clocks']['reasons']}, exact {j['config']['exactness']}")
except Exception as e:
    print(f"N={n}: FAILED {e}", open(f"gpurun_out/scale_n{n}.err").read()[-600:])
PY
done
