#!/bin/bash
# round 2: full GPU suite + captures + secondary benches (tag = $1)
T=${1:-r2e}
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
bash tools/gpu/r2_profiles.sh ${T}p
timeout 900 python tools/ratio_table.py 64 > gpurun_out/${T}_ratio.md 2> gpurun_out/${T}_ratio.err; tail -3 gpurun_out/${T}_ratio.md
timeout 1200 python bench.py --config c3 --steps 2 --warmup 1 > gpurun_out/${T}_bench_c3.log 2>&1; tail -1 gpurun_out/${T}_bench_c3.log | cut -c1-300
timeout 900 python bench.py --config c4 > gpurun_out/${T}_bench_c4.log 2>&1; tail -1 gpurun_out/${T}_bench_c4.log | cut -c1-300
timeout 600 python bench.py --config c2 > gpurun_out/${T}_bench_c2.log 2>&1; tail -1 gpurun_out/${T}_bench_c2.log | cut -c1-300
