#!/bin/bash
# quick check of the last kernel change (near-period sources in M0) on the device (tag = $1)
T=${1:-r2y}
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py -m gpu -x -q -k "deflate or history or c5 or c2 or write_roundtrip" > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
timeout 300 python bench.py --size-gib 4 --level 1 --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench_l1.log 2>&1; tail -1 gpurun_out/${T}_bench_l1.log | cut -c1-200
python tools/kernel_smoke.py > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
