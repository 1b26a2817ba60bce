#!/bin/bash
# round 2, session 2: read path / zip writer iteration (tag = $1)
T=${1:-r2d}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_dropin_cli.py -m gpu -x -q --durations=8 > gpurun_out/${T}_pytest.log 2>&1; tail -12 gpurun_out/${T}_pytest.log
for c in c3 c4; do
  timeout 1200 python bench.py --config $c --no-cpu > gpurun_out/${T}_bench_$c.log 2> gpurun_out/${T}_bench_$c.err; tail -1 gpurun_out/${T}_bench_$c.log | cut -c1-300
done
MZ_CUDA_READ_STATS=1 timeout 900 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu > /dev/null 2> gpurun_out/${T}_c3_readstats.log; grep "read side" gpurun_out/${T}_c3_readstats.log | tail -2
MZ_CUDA_READ_AHEAD=0 MZ_CUDA_READ_STATS=1 timeout 900 python bench.py --config c3 --steps 2 --warmup 1 --no-cpu 2> gpurun_out/${T}_c3_noahead.err | tail -1 | cut -c1-200
du -sh gpurun_out
