#!/bin/bash
# round 2: multi-GPU checks on N GPUs of one box (N = $1, tag = $2)
N=${1:-2}; T=${2:-r2d}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${T}_pytest_multi.log 2>&1; tail -3 gpurun_out/${T}_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${T}_bench_n$N.log 2>&1; tail -1 gpurun_out/${T}_bench_n$N.log | cut -c1-500
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench_n1.log 2>&1; tail -1 gpurun_out/${T}_bench_n1.log | cut -c1-300
