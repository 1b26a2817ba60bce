#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
MZ_CUDA_TRACE=1 timeout 600 python tools/bench_inflate.py long 1024 > gpurun_out/trace.log 2>&1
grep "K6 kernels" gpurun_out/trace.log | head -4; grep GBps gpurun_out/trace.log
timeout 900 python tools/bench_inflate.py > gpurun_out/bench_inflate.log 2>&1; grep -v "^mz_" gpurun_out/bench_inflate.log
mkdir -p /dev/shm/zb && cd /dev/shm/zb
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda c4.zip 100000 65536 6 cuda > /root/repo/gpurun_out/zipbatch.log 2>&1
rm -f c4.zip; timeout 600 /root/repo/oracle/_ref/zipbatch_cuda c4.zip 100000 65536 6 cuda >> /root/repo/gpurun_out/zipbatch.log 2>&1
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda c4l1.zip 100000 65536 1 cuda >> /root/repo/gpurun_out/zipbatch.log 2>&1
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda r.zip 4000 65536 6 ref >> /root/repo/gpurun_out/zipbatch.log 2>&1
python -c "
import zipfile,time
t=time.time(); z=zipfile.ZipFile('c4.zip'); n=len(z.namelist()); bad=z.testzip(); print('zipfile check', n, bad, round(time.time()-t,1),'s')" >> /root/repo/gpurun_out/zipbatch.log 2>&1
cat /root/repo/gpurun_out/zipbatch.log | cut -c1-420
rm -rf /dev/shm/zb
cd /root/repo
timeout 900 python bench.py > gpurun_out/bench_16g.log 2>&1; echo "bench exit $?" >> gpurun_out/bench_16g.log; tail -2 gpurun_out/bench_16g.log | cut -c1-2400
