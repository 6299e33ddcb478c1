set -x
timeout 900 python -m pytest tests/test_gpu_retrieval.py -x -q -m gpu 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_n2_two_phase.json 2> gpurun_out/bench_n2_two_phase.err; tail -c 600 gpurun_out/bench_n2_two_phase.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --single-phase --no-secondary --no-cpu-baseline > gpurun_out/bench_n2_single_phase.json 2>/dev/null
