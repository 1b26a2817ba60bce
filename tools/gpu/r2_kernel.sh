#!/bin/bash
# round 2: deflate kernel iteration -- parity subset, bench at three levels, ncu full capture (tag = $1)
T=${1:-r2b}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "parity or configs" > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
for lv in 1 2 6; do timeout 300 python bench.py --size-gib 4 --level $lv --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench_l$lv.log 2>&1; tail -1 gpurun_out/${T}_bench_l$lv.log | cut -c1-300; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deflate_chunks -s 1 -c 1 -o gpurun_out/${T}_prof_deflate python bench.py --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_ncu_deflate.log 2>&1
ls -la gpurun_out/${T}*.ncu-rep
