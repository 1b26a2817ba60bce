#!/bin/bash
# round 2, session 2: the history variant (levels 6-9) on the device (tag = $1): parity, level-6 rates, C2 / C4, ratio table
T=${1:-r2e}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_dropin_cli.py tests/test_gpu_fullsize.py -m gpu -x -q -k "not four_gib and not hundred" > gpurun_out/${T}_pytest.log 2>&1; tail -4 gpurun_out/${T}_pytest.log
for lv in 1 4 6; do timeout 300 python bench.py --size-gib 4 --level $lv --steps 4 --warmup 3 --no-e2e --no-cpu > gpurun_out/${T}_bench_l$lv.log 2>&1; tail -1 gpurun_out/${T}_bench_l$lv.log | python -c "
import sys,json
j=json.loads(sys.stdin.readline()); print('L$lv', j['value'], j['unit'], 'kernel', j['roofline']['achieved'], 'GB/s ratio', j.get('ratio'))"; done
for c in c2 c4; do
  timeout 1200 python bench.py --config $c --no-cpu > gpurun_out/${T}_bench_$c.log 2> gpurun_out/${T}_bench_$c.err; tail -1 gpurun_out/${T}_bench_$c.log | cut -c1-300
done
timeout 900 python tools/ratio_table.py 64 > gpurun_out/${T}_ratio.md 2> gpurun_out/${T}_ratio.err; cat gpurun_out/${T}_ratio.md | cut -c1-200
du -sh gpurun_out
