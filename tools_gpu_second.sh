#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --size-gib 2 --steps 3 --warmup 3 --cpu-sample-mib 1024 > gpurun_out/bench_2g.log 2>&1; echo "bench2g exit $?" >> gpurun_out/bench_2g.log
timeout 900 python bench.py > gpurun_out/bench_16g.log 2>&1; echo "bench16g exit $?" >> gpurun_out/bench_16g.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1; echo "benchref exit $?" >> gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --size-gib 1 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:deflate_chunks -s 1 -c 1 -o gpurun_out/prof_deflate python bench.py --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_deflate.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:crc32_segments -s 1 -c 1 -o gpurun_out/prof_crc python bench.py --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_crc.log 2>&1
tail -2 gpurun_out/bench_2g.log; tail -2 gpurun_out/bench_16g.log; tail -2 gpurun_out/bench_ref.log; ls -la gpurun_out
