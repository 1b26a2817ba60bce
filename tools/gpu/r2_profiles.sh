#!/bin/bash
# round 2: one ncu --set full capture per kernel (tag = $1); summaries are made afterwards with tools/ncu_summary.py
T=${1:-r2p}
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
timeout 600 $N -k regex:deflate_chunks -s 1 -c 1 -o gpurun_out/${T}_deflate python bench.py --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_deflate.log 2>&1
timeout 600 $N -k regex:deflate_chunks -s 1 -c 1 -o gpurun_out/${T}_deflate_l6 python bench.py --size-gib 1 --level 6 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_deflate_l6.log 2>&1
timeout 600 $N -k regex:crc32_segments -s 2 -c 1 -o gpurun_out/${T}_crc python bench.py --size-gib 4 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_crc.log 2>&1
timeout 600 $N -k regex:gather_slots -s 1 -c 1 -o gpurun_out/${T}_gather python bench.py --size-gib 4 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/${T}_gather.log 2>&1
for k in find scan emit; do timeout 900 $N -k regex:inflate_spec_$k -s 1 -c 1 -o gpurun_out/${T}_k6$k python tools/bench_inflate.py long 512 > gpurun_out/${T}_k6$k.log 2>&1; done
timeout 600 $N -k regex:inflate_streams -c 1 -o gpurun_out/${T}_k5batch python tools/bench_inflate.py batch 8192 > gpurun_out/${T}_k5batch.log 2>&1
timeout 600 $N -k regex:sha256_batch -s 3 -c 1 -o gpurun_out/${T}_sha python tools/bench_sha.py 20000 65536 > gpurun_out/${T}_sha.log 2>&1
timeout 300 python tools/bench_sha.py 100000 65536 > gpurun_out/${T}_sha_bench.log 2>&1; tail -1 gpurun_out/${T}_sha_bench.log
timeout 900 python tools/bench_inflate.py > gpurun_out/${T}_inflate_bench.jsonl 2>&1; tail -4 gpurun_out/${T}_inflate_bench.jsonl | cut -c1-250
ls -la gpurun_out/${T}*.ncu-rep | awk '{print $5, $9}'
