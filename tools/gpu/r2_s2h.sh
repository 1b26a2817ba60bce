#!/bin/bash
# the long-member mismatch with the copy verification on (tag = $1)
T=${1:-r2h}
mkdir -p gpurun_out
MZ_CUDA_VERIFY_COPY=1 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -s > gpurun_out/${T}_verify.log 2>&1; echo "verify: $(tail -1 gpurun_out/${T}_verify.log)"; grep -c VERIFY_COPY gpurun_out/${T}_verify.log; grep VERIFY_COPY gpurun_out/${T}_verify.log | head -20
