#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "c3 or c4 or inflate or read or dropin or decompress or zip_batch or golden or foreign" > gpurun_out/pytest_gpu_inf.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_inf.log; tail -4 gpurun_out/pytest_gpu_inf.log
timeout 900 python tools/bench_inflate.py > gpurun_out/bench_inflate.log 2>&1; grep -v "^mz_" gpurun_out/bench_inflate.log
MZ_CUDA_TRACE=1 timeout 600 python tools/bench_inflate.py long 1024 > gpurun_out/trace.log 2>&1
grep "K6 kernels" gpurun_out/trace.log | head -4
