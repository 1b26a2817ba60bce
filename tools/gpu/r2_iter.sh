#!/bin/bash
# round 2: one iteration -- parity subset, deflate bench at three levels, one full ncu capture boiled down on the box,
# the C3 read path with the caller-thread timers, C4 with its stage times (tag = $1)
T=${1:-r2f}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "parity or configs or dropin" > gpurun_out/${T}_pytest.log 2>&1; tail -2 gpurun_out/${T}_pytest.log
for lv in 1 2 6; do timeout 300 python bench.py --size-gib 4 --level $lv --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench_l$lv.log 2>&1; tail -1 gpurun_out/${T}_bench_l$lv.log | python -c "
import sys,json
j=json.loads(sys.stdin.readline()); print('L$lv', j['value'], j['unit'], 'kernel', j['roofline']['achieved'], 'GB/s ratio', j.get('ratio'))"; done
N="ncu --set full --clock-control none --import-source on"
timeout 900 $N -k regex:deflate_chunks -s 1 -c 1 -f -o gpurun_out/${T}_deflate python bench.py --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_ncu_deflate.log 2>&1
if [ -f gpurun_out/${T}_deflate.ncu-rep ]; then
  python tools/ncu_summary.py gpurun_out/${T}_deflate.ncu-rep $((1<<30)) > gpurun_out/${T}_deflate.summary.txt 2>&1
  ncu -i gpurun_out/${T}_deflate.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${T}_deflate.source.csv 2>/dev/null
  python tools/ncu_line_ops.py gpurun_out/${T}_deflate.source.csv 70 > gpurun_out/${T}_deflate.lines.txt 2>&1
  python tools/ncu_segments.py gpurun_out/${T}_deflate.ncu-rep > gpurun_out/${T}_deflate.segments.txt 2>&1
  head -12 gpurun_out/${T}_deflate.summary.txt
fi
MZ_CUDA_READ_STATS=1 timeout 900 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu > gpurun_out/${T}_bench_c3.log 2> gpurun_out/${T}_c3_trace.log; tail -1 gpurun_out/${T}_bench_c3.log | cut -c1-200; grep "read side" gpurun_out/${T}_c3_trace.log | tail -3
timeout 900 python bench.py --config c4 --no-cpu > gpurun_out/${T}_bench_c4.log 2>&1; tail -1 gpurun_out/${T}_bench_c4.log | python -c "
import sys,json
j=json.loads(sys.stdin.readline()); print('C4', j['value'], j['unit'], j['entries_per_s'], 'entries/s', j['e2e'])"
du -sh gpurun_out
