#!/bin/bash
# round 2: the 8-GPU box (gpurun --gpus 8; charged 8x, so every command is short): multi-GPU tests, C5 strong scaling at
# N = 8, 4, 2, 1 as the driver launches it, C4 with the entries sharded over all GPUs inside one archive writer (tag = $1)
T=${1:-r2h}
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/${T}_ngpu.txt
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/${T}_pytest_multi.log 2>&1; tail -2 gpurun_out/${T}_pytest_multi.log
for N in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 8 --warmup 3 --no-cpu > gpurun_out/${T}_bench_n$N.log 2> gpurun_out/${T}_bench_n$N.err
  tail -1 gpurun_out/${T}_bench_n$N.log | python -c "
import sys,json
try:
    j=json.loads(sys.stdin.readline()); print('N=$N', j['value'], j['unit'], 'kernel', j['roofline']['achieved'], 'e2e', j['e2e']['value'])
except Exception as e: print('N=$N failed', e)"
done
timeout 300 python bench.py --gpus 1 --steps 8 --warmup 3 --no-cpu > gpurun_out/${T}_bench_n1.log 2> gpurun_out/${T}_bench_n1.err; tail -1 gpurun_out/${T}_bench_n1.log | cut -c1-200
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --config c4 --gpus 8 --no-cpu > gpurun_out/${T}_bench_c4_n8.log 2> gpurun_out/${T}_bench_c4_n8.err; tail -1 gpurun_out/${T}_bench_c4_n8.log | cut -c1-300
