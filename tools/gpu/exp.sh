#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/bench_inflate.py batch 8192 > gpurun_out/exp_batch.log 2>&1; tail -1 gpurun_out/exp_batch.log
MZ_CUDA_TRACE=1 timeout 600 python tools/bench_inflate.py long 1024 > gpurun_out/exp_trace.log 2>&1
grep "K6 kernels" gpurun_out/exp_trace.log | head -4; grep GBps gpurun_out/exp_trace.log
timeout 900 python -m pytest tests -m gpu -x -q -k "c3 or c4 or inflate or zip_batch or threads" > gpurun_out/exp_pytest.log 2>&1; tail -3 gpurun_out/exp_pytest.log
