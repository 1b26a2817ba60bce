#!/bin/bash
# quick check of the last kernel change (near-period sources in M0) on the device (tag = $1)
T=${1:-r2y}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "deflate_batch or history or nonfinal" > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
for i in 1 2; do timeout 300 python bench.py --size-gib 4 --level 1 --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench_l1_$i.log 2>&1; tail -1 gpurun_out/${T}_bench_l1_$i.log | python -c "
import sys,json
j=json.loads(sys.stdin.readline()); print('L1', j['value'], 'kernel', j['roofline']['achieved'], j['roofline']['ms_per_launch'])"; done
python tools/kernel_smoke.py > gpurun_out/${T}_smoke.log 2>&1; tail -1 gpurun_out/${T}_smoke.log
