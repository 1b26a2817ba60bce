#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_16g.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_16g.log; tail -2 gpurun_out/bench_16g.log | cut -c1-300
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-300
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deflate_chunks -s 1 -c 1 -o gpurun_out/prof_deflate python bench.py --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_deflate.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:inflate_spec_scan -c 1 -o gpurun_out/prof_k6scan python tools/bench_inflate.py long 256 > gpurun_out/ncu_k6scan.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:crc32_segments -s 2 -c 1 -o gpurun_out/prof_crc python bench.py --size-gib 4 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_crc.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/launches.csv | cut -c1-200
