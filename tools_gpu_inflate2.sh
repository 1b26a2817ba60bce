#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "c3 or c4 or inflate or read or dropin or decompress" > gpurun_out/pytest_gpu_inf.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_inf.log
timeout 900 python tools/bench_inflate.py > gpurun_out/bench_inflate.log 2>&1; echo "inflate exit $?" >> gpurun_out/bench_inflate.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inflate_streams -s 1 -c 1 -o gpurun_out/prof_inflate python tools/bench_inflate.py single 4 > gpurun_out/ncu_inflate.log 2>&1
tail -5 gpurun_out/pytest_gpu_inf.log; cat gpurun_out/bench_inflate.log; tail -3 gpurun_out/ncu_inflate.log
