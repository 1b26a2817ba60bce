#!/bin/bash
mkdir -p gpurun_out
run() { # nproc port extra...
  n=$1; port=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 3 --warmup 3 --no-e2e "$@" 2>&1 | grep '^{' | tail -1
}
run 8 29531 > gpurun_out/bench_n8.log
run 8 29532 --sub-batches 8 > gpurun_out/bench_n8_nb8.log
run 8 29533 --sub-batches 1 > gpurun_out/bench_n8_nb1.log
run 4 29534 > gpurun_out/bench_n4.log
run 8 29535 --impl reference > gpurun_out/bench_n8_ref.log
for f in bench_n8 bench_n8_nb8 bench_n8_nb1 bench_n4 bench_n8_ref; do echo $f; cut -c1-260 gpurun_out/$f.log; done
