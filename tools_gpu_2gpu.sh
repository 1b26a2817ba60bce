#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --size-gib 4 --steps 3 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/bench_n1_4g.log 2>&1; echo "bench1 exit $?" >> gpurun_out/bench_n1_4g.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.log 2>&1; echo "bench2 exit $?" >> gpurun_out/bench_n2.log
timeout 600 python tools/bench_inflate.py > gpurun_out/bench_inflate.log 2>&1; echo "inflate exit $?" >> gpurun_out/bench_inflate.log
tail -2 gpurun_out/bench_n1_4g.log | cut -c1-600; tail -3 gpurun_out/bench_n2.log | cut -c1-900; cat gpurun_out/bench_inflate.log
