#!/bin/bash
# round 2, session 2, last single-GPU run (tag = $1): the whole GPU suite, the default bench with its reference arm, the launch list
T=${1:-r2z}
mkdir -p gpurun_out
( time timeout 1300 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/${T}_pytest.log 2>&1; tail -6 gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench_c5.log 2> gpurun_out/${T}_bench_c5.err; tail -1 gpurun_out/${T}_bench_c5.log | cut -c1-300
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_c5_ref.log 2> gpurun_out/${T}_bench_c5_ref.err; tail -1 gpurun_out/${T}_bench_c5_ref.log | cut -c1-200
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --size-gib 4 --steps 2 --warmup 1 --no-cpu > gpurun_out/${T}_launches.log 2>&1
timeout 300 python bench.py --config c1 --no-cpu > gpurun_out/${T}_bench_c1.log 2>/dev/null; tail -1 gpurun_out/${T}_bench_c1.log | cut -c1-160
du -sh gpurun_out
