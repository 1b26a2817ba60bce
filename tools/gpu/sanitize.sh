#!/bin/bash
# device-side memory checking of every kernel (compute-sanitizer memcheck; racecheck is not used: the deflate
# kernel's hash table is racy by design)
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/memcheck.log python tools/kernel_smoke.py > gpurun_out/memcheck_run.log 2>&1; echo "memcheck exit $?" >> gpurun_out/memcheck_run.log
tail -4 gpurun_out/memcheck_run.log; tail -6 gpurun_out/memcheck.log
cd /tmp && mkdir -p zs && cd zs
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file /root/repo/gpurun_out/memcheck_zip.log /root/repo/oracle/_ref/zipbatch_cuda s.zip 300 65536 6 cuda > /root/repo/gpurun_out/memcheck_zip_run.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file /root/repo/gpurun_out/memcheck_zipx.log /root/repo/oracle/_ref/zipbatch_cuda s.zip 300 65536 6 extract >> /root/repo/gpurun_out/memcheck_zip_run.log 2>&1
cut -c1-200 /root/repo/gpurun_out/memcheck_zip_run.log; tail -n 3 /root/repo/gpurun_out/memcheck_zip.log; tail -n 3 /root/repo/gpurun_out/memcheck_zipx.log
