#!/bin/bash
# round 2, session 2, first run (tag = $1): the whole GPU suite with per-test durations, the default bench (both arms),
# the deflate kernel's full ncu capture boiled down on the box, the launch list of the bench command, the ratio table.
T=${1:-r2a}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,memory.total --format=csv > gpurun_out/${T}_gpu.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q --durations=40 ) > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
timeout 900 python bench.py > gpurun_out/${T}_bench_c5.log 2> gpurun_out/${T}_bench_c5.err; tail -1 gpurun_out/${T}_bench_c5.log | cut -c1-400
timeout 900 python bench.py --impl reference > gpurun_out/${T}_bench_c5_ref.log 2> gpurun_out/${T}_bench_c5_ref.err; tail -1 gpurun_out/${T}_bench_c5_ref.log | cut -c1-300
N="ncu --set full --clock-control none --import-source on"
for lv in 1 6; do
  timeout 900 $N -k regex:deflate_chunks -s 1 -c 1 -f -o gpurun_out/${T}_deflate_l$lv python bench.py --size-gib 1 --level $lv --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_ncu_deflate_l$lv.log 2>&1
  R=gpurun_out/${T}_deflate_l$lv
  if [ -f $R.ncu-rep ]; then
    python tools/ncu_summary.py $R.ncu-rep $((1<<30)) > $R.summary.txt 2>&1
    ncu -i $R.ncu-rep --page raw --csv > $R.raw.csv 2>/dev/null
    ncu -i $R.ncu-rep --page source --csv --print-source cuda,sass > $R.source.csv 2>/dev/null
    python tools/ncu_line_ops.py $R.source.csv 70 > $R.lines.txt 2>&1
    python tools/ncu_segments.py $R.ncu-rep > $R.segments.txt 2>&1
    head -12 $R.summary.txt
    [ $lv = 6 ] && rm -f $R.ncu-rep $R.source.csv
  fi
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --size-gib 4 --steps 2 --warmup 1 --no-cpu > gpurun_out/${T}_launches.log 2>&1
timeout 900 python tools/ratio_table.py 64 > gpurun_out/${T}_ratio.md 2> gpurun_out/${T}_ratio.err; tail -6 gpurun_out/${T}_ratio.md
du -sh gpurun_out
