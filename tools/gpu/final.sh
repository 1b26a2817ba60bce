#!/bin/bash
# what the driver runs at round end, in one call: the GPU suite, smoke(), both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench_16g.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_16g.log; tail -2 gpurun_out/bench_16g.log | cut -c1-600
