#!/bin/bash
# last check of the session on the device: the M0 near-source rule after the period-3 fix, run thinning (tag = $1)
T=${1:-r2w}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "runs_and_near or history or nonfinal or (deflate_batch and text)" > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
timeout 200 python bench.py --size-gib 4 --level 1 --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench_l1.log 2>&1; tail -1 gpurun_out/${T}_bench_l1.log | python -c "
import sys,json
j=json.loads(sys.stdin.readline()); print('L1', j['value'], 'kernel', j['roofline']['achieved'], j['roofline']['ms_per_launch'])"
