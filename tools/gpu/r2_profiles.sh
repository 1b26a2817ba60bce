#!/bin/bash
# round 2: one ncu --set full capture per kernel (tag = $1). gpurun brings back at most 64 MiB, so each report is boiled down ON
# THE BOX (tools/ncu_summary.py + the raw page as csv; for the deflate kernel also ncu's source/SASS correlation) and only the
# deflate report itself is kept.
T=${1:-r2p}
mkdir -p gpurun_out
N="ncu --set full --clock-control none --import-source on"
cap() { # name, kernel regex, skip, algorithmic bytes, command...
  local name=$1 rx=$2 skip=$3 alg=$4; shift 4
  timeout 900 $N -k regex:$rx -s $skip -c 1 -f -o gpurun_out/${T}_$name "$@" > gpurun_out/${T}_$name.log 2>&1
  if [ -f gpurun_out/${T}_$name.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/${T}_$name.ncu-rep $alg > gpurun_out/${T}_$name.summary.txt 2>&1
    ncu -i gpurun_out/${T}_$name.ncu-rep --page raw --csv > gpurun_out/${T}_$name.raw.csv 2>/dev/null
    case $name in deflate*) ncu -i gpurun_out/${T}_$name.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/${T}_$name.source.csv 2>/dev/null
                            python tools/ncu_line_ops.py gpurun_out/${T}_$name.source.csv 60 > gpurun_out/${T}_$name.lines.txt 2>&1
                            python tools/ncu_segments.py gpurun_out/${T}_$name.ncu-rep > gpurun_out/${T}_$name.segments.txt 2>&1 ;; esac
    case $name in deflate) ;; *) rm -f gpurun_out/${T}_$name.ncu-rep ;; esac
    head -8 gpurun_out/${T}_$name.summary.txt | tail -5
  else echo "capture $name FAILED"; tail -5 gpurun_out/${T}_$name.log; fi
}
G1=$((1<<30)); G4=$((4<<30))
cap deflate deflate_chunks 1 $G1 python bench.py --size-gib 1 --steps 1 --warmup 1 --no-e2e --no-cpu
cap deflate_l6 deflate_chunks 1 $G1 python bench.py --size-gib 1 --level 6 --steps 1 --warmup 1 --no-e2e --no-cpu
cap crc crc32_segments 1 $G4 python bench.py --size-gib 4 --steps 1 --warmup 1 --no-e2e --no-cpu
cap gather gather_slots 1 0 python bench.py --size-gib 4 --steps 1 --warmup 1 --no-e2e --no-cpu
for k in find scan emit; do cap k6$k inflate_spec_$k 1 0 python tools/bench_inflate.py long 512; done
cap k5batch inflate_streams 0 0 python tools/bench_inflate.py batch 8192
cap sha sha256_batch 3 $((20000*65536)) python tools/bench_sha.py 20000 65536
timeout 300 python tools/bench_sha.py 100000 65536 > gpurun_out/${T}_sha_bench.log 2>&1; tail -1 gpurun_out/${T}_sha_bench.log
timeout 900 python tools/bench_inflate.py > gpurun_out/${T}_inflate_bench.jsonl 2>&1; tail -4 gpurun_out/${T}_inflate_bench.jsonl | cut -c1-250
du -sh gpurun_out
