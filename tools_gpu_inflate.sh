#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 600 python tools/bench_inflate.py > gpurun_out/bench_inflate.log 2>&1; echo "inflate exit $?" >> gpurun_out/bench_inflate.log
tail -5 gpurun_out/pytest_gpu.log; cat gpurun_out/bench_inflate.log
