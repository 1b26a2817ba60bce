#!/bin/bash
mkdir -p gpurun_out
MZ_CUDA_TRACE=1 timeout 600 python tools/bench_inflate.py long 1024 > gpurun_out/trace.log 2>&1
grep "K6 round\|K6 kernels" gpurun_out/trace.log | head -6; grep GBps gpurun_out/trace.log
timeout 900 python -m pytest tests -m gpu -x -q -k "c3 or c4 or inflate or read or dropin or decompress or zip_batch" > gpurun_out/pytest_gpu_inf.log 2>&1; tail -3 gpurun_out/pytest_gpu_inf.log
cd /tmp && mkdir -p zb && cd zb
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda c4.zip 100000 65536 6 cuda > /root/repo/gpurun_out/zipbatch.log 2>&1
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda c4.zip 100000 65536 6 cuda >> /root/repo/gpurun_out/zipbatch.log 2>&1
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda c4l1.zip 100000 65536 1 cuda >> /root/repo/gpurun_out/zipbatch.log 2>&1
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda r.zip 4000 65536 6 ref >> /root/repo/gpurun_out/zipbatch.log 2>&1
python -c "
import zipfile,time
t=time.time(); z=zipfile.ZipFile('c4.zip'); n=len(z.namelist()); bad=z.testzip(); print('zipfile check', n, bad, round(time.time()-t,1),'s')" >> /root/repo/gpurun_out/zipbatch.log 2>&1
cat /root/repo/gpurun_out/zipbatch.log
