#!/bin/bash
# round 2, session 2: one deflate-kernel iteration (tag = $1): parity subset, bench at levels 1 and 6 (4 GiB), the default
# bench, one ncu --set full capture of the level-1 kernel boiled down on the box
T=${1:-r2c}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
for lv in 1 6; do timeout 300 python bench.py --size-gib 4 --level $lv --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench_l$lv.log 2>&1; tail -1 gpurun_out/${T}_bench_l$lv.log | python -c "
import sys,json
j=json.loads(sys.stdin.readline()); print('L$lv', j['value'], j['unit'], 'kernel', j['roofline']['achieved'], 'GB/s ratio', j.get('ratio'))"; done
timeout 600 python bench.py --no-cpu > gpurun_out/${T}_bench_c5.log 2> gpurun_out/${T}_bench_c5.err; tail -1 gpurun_out/${T}_bench_c5.log | cut -c1-200
N="ncu --set full --clock-control none --import-source on"
R=gpurun_out/${T}_deflate_l1
timeout 900 $N -k regex:deflate_chunks -s 1 -c 1 -f -o $R python bench.py --size-gib 1 --level 1 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_ncu_deflate_l1.log 2>&1
if [ -f $R.ncu-rep ]; then
  python tools/ncu_summary.py $R.ncu-rep $((1<<30)) > $R.summary.txt 2>&1
  ncu -i $R.ncu-rep --page raw --csv > $R.raw.csv 2>/dev/null
  ncu -i $R.ncu-rep --page source --csv --print-source cuda,sass > $R.source.csv 2>/dev/null
  python tools/ncu_line_ops.py $R.source.csv 70 > $R.lines.txt 2>&1
  python tools/ncu_segments.py $R.ncu-rep > $R.segments.txt 2>&1
  head -12 $R.summary.txt
  rm -f $R.ncu-rep
fi
du -sh gpurun_out
