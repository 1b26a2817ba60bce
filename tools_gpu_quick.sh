#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "deflate or stream_write or roundtrip or written" > gpurun_out/pytest_gpu_quick.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_quick.log
timeout 600 python bench.py --size-gib 4 --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_4g.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_4g.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deflate_chunks -s 1 -c 1 -o gpurun_out/prof_deflate python bench.py --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_deflate.log 2>&1
tail -3 gpurun_out/pytest_gpu_quick.log; tail -2 gpurun_out/bench_4g.log | cut -c1-1500
