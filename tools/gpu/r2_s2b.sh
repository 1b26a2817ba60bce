#!/bin/bash
# round 2, session 2, second run (tag = $1): bench configs c1..c4 with their reference arms, the C3 read-side timers, one
# ncu --set full capture of each kernel other than deflate (boiled down on the box), the inflate / SHA micro-benchmarks.
T=${1:-r2b}
mkdir -p gpurun_out
for c in c1 c2 c3 c4; do
  timeout 1200 python bench.py --config $c > gpurun_out/${T}_bench_$c.log 2> gpurun_out/${T}_bench_$c.err; tail -1 gpurun_out/${T}_bench_$c.log | cut -c1-300
  timeout 900 python bench.py --config $c --impl reference > gpurun_out/${T}_bench_${c}_ref.log 2> gpurun_out/${T}_bench_${c}_ref.err; tail -1 gpurun_out/${T}_bench_${c}_ref.log | cut -c1-200
done
MZ_CUDA_READ_STATS=1 timeout 900 python bench.py --config c3 --steps 1 --warmup 1 --no-cpu > /dev/null 2> gpurun_out/${T}_c3_readstats.log; grep "read side" gpurun_out/${T}_c3_readstats.log | tail -2
N="ncu --set full --clock-control none --import-source on"
cap() { # name, kernel regex, skip, algorithmic bytes, command...
  local name=$1 rx=$2 skip=$3 alg=$4; shift 4
  timeout 600 $N -k regex:$rx -s $skip -c 1 -f -o gpurun_out/${T}_$name "$@" > gpurun_out/${T}_$name.log 2>&1
  if [ -f gpurun_out/${T}_$name.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/${T}_$name.ncu-rep $alg > gpurun_out/${T}_$name.summary.txt 2>&1
    ncu -i gpurun_out/${T}_$name.ncu-rep --page raw --csv > gpurun_out/${T}_$name.raw.csv 2>/dev/null
    rm -f gpurun_out/${T}_$name.ncu-rep
    head -8 gpurun_out/${T}_$name.summary.txt | tail -5
  else echo "capture $name FAILED"; tail -5 gpurun_out/${T}_$name.log; fi
}
G4=$((4<<30))
cap crc crc32_segments 1 $G4 python bench.py --size-gib 4 --steps 1 --warmup 1 --no-e2e --no-cpu
cap gather gather_slots 1 0 python bench.py --size-gib 4 --steps 1 --warmup 1 --no-e2e --no-cpu
for k in find scan emit; do cap k6$k inflate_spec_$k 1 0 python tools/bench_inflate.py long 512; done
cap k5batch inflate_streams 0 0 python tools/bench_inflate.py batch 8192
cap sha sha256_batch 3 $((20000*65536)) python tools/bench_sha.py 20000 65536
timeout 300 python tools/bench_sha.py 100000 65536 > gpurun_out/${T}_sha_bench.log 2>&1; tail -1 gpurun_out/${T}_sha_bench.log
timeout 600 python tools/bench_inflate.py > gpurun_out/${T}_inflate_bench.jsonl 2>&1; tail -4 gpurun_out/${T}_inflate_bench.jsonl | cut -c1-250
du -sh gpurun_out
