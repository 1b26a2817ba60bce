#!/bin/bash
# round 2, the run the committed profiles/ files come from (tag = $1): the whole GPU suite, one ncu --set full capture per kernel
# (boiled down on the box), the launch list of the default bench command, the ratio table, every bench config with its
# reference arm. gpurun brings back at most 64 MiB.
T=${1:-r2g}
mkdir -p gpurun_out
timeout 2700 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log
bash tools/gpu/r2_profiles.sh ${T}p
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --size-gib 4 --steps 2 --warmup 1 --no-cpu > gpurun_out/${T}_launches.log 2>&1
timeout 900 python tools/ratio_table.py 64 > gpurun_out/${T}_ratio.md 2> gpurun_out/${T}_ratio.err; tail -4 gpurun_out/${T}_ratio.md
for c in c5 c1 c2 c3 c4; do
  timeout 1500 python bench.py --config $c > gpurun_out/${T}_bench_$c.log 2> gpurun_out/${T}_bench_$c.err; tail -1 gpurun_out/${T}_bench_$c.log | cut -c1-260
  timeout 900 python bench.py --config $c --impl reference > gpurun_out/${T}_bench_${c}_ref.log 2> gpurun_out/${T}_bench_${c}_ref.err; tail -1 gpurun_out/${T}_bench_${c}_ref.log | cut -c1-200
done
MZ_CUDA_READ_STATS=1 timeout 900 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu > /dev/null 2> gpurun_out/${T}_c3_readstats.log; grep "read side" gpurun_out/${T}_c3_readstats.log | tail -2
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/${T}_gpu.txt
du -sh gpurun_out
