#!/bin/bash
mkdir -p gpurun_out
MZ_CUDA_TRACE=1 timeout 600 python - > gpurun_out/trace.log 2>&1 <<'PY'
import sys, os, zlib, time, json
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import cuharness
p = cuharness.pkg(); lib = p.load(); lib.mz_cuda_init()
tl = cuharness.TestLib()
n = 256 << 20
host = bytes(p.textgen(n, seed=9).cpu().numpy().tobytes())
co = zlib.compressobj(6, zlib.DEFLATED, 31); comp = co.compress(host) + co.flush()
for rep in range(2):
    t0 = time.perf_counter()
    out, info = tl.decompress(lib.mz_stream_cuda_create, comp, n, window_bits=31, read_size=1 << 20)
    dt = time.perf_counter() - t0
    print(json.dumps({"s": dt, "GBps": n / dt / 1e9, "ok": zlib.crc32(out) == zlib.crc32(host), "info": info}), flush=True)
PY
grep -c "K6 round" gpurun_out/trace.log; grep -c "K5 launch" gpurun_out/trace.log; grep "K6 round\|K6 kernels" gpurun_out/trace.log | head -12; grep GBps gpurun_out/trace.log
timeout 300 python tools/bench_inflate.py single 16 > gpurun_out/single.log 2>&1; tail -1 gpurun_out/single.log
timeout 300 python tools/bench_inflate.py batch 8192 > gpurun_out/batch.log 2>&1; tail -1 gpurun_out/batch.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:inflate_streams -s 1 -c 1 -o gpurun_out/prof_inflate_batch python tools/bench_inflate.py batch 8192 > gpurun_out/ncu_inflate_batch.log 2>&1
tail -2 gpurun_out/ncu_inflate_batch.log
