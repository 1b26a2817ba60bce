#!/bin/bash
# round 2, first calibration of the v3 deflate kernel: parity tests, bench, ncu full capture with source
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "parity or configs" > gpurun_out/r2a_pytest.log 2>&1; tail -3 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --no-e2e --no-cpu > gpurun_out/r2a_bench.log 2>&1; tail -1 gpurun_out/r2a_bench.log | cut -c1-600
for lv in 1 2 6; do timeout 300 python bench.py --size-gib 2 --level $lv --steps 3 --warmup 2 --no-e2e --no-cpu > gpurun_out/r2a_bench_l$lv.log 2>&1; tail -1 gpurun_out/r2a_bench_l$lv.log | cut -c1-400; done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deflate_chunks -s 1 -c 1 -o gpurun_out/r2a_prof_deflate python bench.py --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2a_ncu_deflate.log 2>&1
ls -la gpurun_out/*.ncu-rep
