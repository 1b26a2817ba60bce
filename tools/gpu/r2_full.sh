#!/bin/bash
# round 2: whole GPU suite (incl. full-size configs) + every bench configuration (tag = $1)
T=${1:-r2c}
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
for c in c1 c2 c4; do timeout 900 python bench.py --config $c > gpurun_out/${T}_bench_$c.log 2>&1; tail -1 gpurun_out/${T}_bench_$c.log | cut -c1-400; done
timeout 1200 python bench.py --config c3 --steps 2 --warmup 1 > gpurun_out/${T}_bench_c3.log 2>&1; tail -1 gpurun_out/${T}_bench_c3.log | cut -c1-400
timeout 900 python bench.py > gpurun_out/${T}_bench_c5.log 2>&1; tail -1 gpurun_out/${T}_bench_c5.log | cut -c1-400
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_c5_ref.log 2>&1; tail -1 gpurun_out/${T}_bench_c5_ref.log | cut -c1-300
