#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e > gpurun_out/bench_n2.log 2>&1; echo "bench2 exit $?" >> gpurun_out/bench_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e --sub-batches 1 > gpurun_out/bench_n2_nb1.log 2>&1; echo "bench2 exit $?" >> gpurun_out/bench_n2_nb1.log
tail -2 gpurun_out/bench_n2.log | cut -c1-400; tail -2 gpurun_out/bench_n2_nb1.log | cut -c1-400
