#!/bin/bash
# reproduce the order of run r2e (parity tests before the config tests), then split (tag = $1)
T=${1:-r2g}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/${T}_order.log 2>&1; echo "parity+configs: $(tail -1 gpurun_out/${T}_order.log)"
if grep -q failed gpurun_out/${T}_order.log; then
  timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "not history" > gpurun_out/${T}_nohist.log 2>&1; echo "without the history test: $(tail -1 gpurun_out/${T}_nohist.log)"
  MZ_CUDA_COPY_THREADS=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q > gpurun_out/${T}_onecopy.log 2>&1; echo "one copy thread: $(tail -1 gpurun_out/${T}_onecopy.log)"
fi
