#!/bin/bash
# bisect the long-member mismatch seen in run r2e (tag = $1)
T=${1:-r2f}
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m pytest tests/test_gpu_configs.py -m gpu -x -q -k "long_member" > gpurun_out/${T}_$name.log 2>&1; echo "$name: $(tail -1 gpurun_out/${T}_$name.log)"
}
run default1 A=1
run default2 A=1
run default3 A=1
run noahead1 MZ_CUDA_READ_AHEAD=0
run noahead2 MZ_CUDA_READ_AHEAD=0
run onecopy1 MZ_CUDA_COPY_THREADS=1
run onecopy2 MZ_CUDA_COPY_THREADS=1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "history or roundtrip or long_member" > gpurun_out/${T}_after_parity.log 2>&1; echo "after parity: $(tail -1 gpurun_out/${T}_after_parity.log)"
