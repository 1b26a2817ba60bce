#!/bin/bash
# first GPU validation: tests, smoke, small bench, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
timeout 600 python bench.py --size-gib 2 --steps 3 --warmup 3 --cpu-sample-mib 512 > gpurun_out/bench_2g.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_2g.log
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/smoke.log; tail -3 gpurun_out/bench_2g.log
