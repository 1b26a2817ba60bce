#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "zip_batch or dropin" > gpurun_out/pytest_gpu_zip.log 2>&1; tail -3 gpurun_out/pytest_gpu_zip.log
cd /tmp && mkdir -p zb && cd zb
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda c4.zip 100000 65536 6 cuda > /root/repo/gpurun_out/zipbatch.log 2>&1
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda c4.zip 100000 65536 6 cuda >> /root/repo/gpurun_out/zipbatch.log 2>&1
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda c4l1.zip 100000 65536 1 cuda >> /root/repo/gpurun_out/zipbatch.log 2>&1
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda r.zip 4000 65536 6 ref >> /root/repo/gpurun_out/zipbatch.log 2>&1
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda r1.zip 8000 65536 1 ref >> /root/repo/gpurun_out/zipbatch.log 2>&1
python -c "
import zipfile,time
t=time.time(); z=zipfile.ZipFile('c4.zip'); n=len(z.namelist()); bad=z.testzip(); print('zipfile check', n, bad, round(time.time()-t,1),'s')" >> /root/repo/gpurun_out/zipbatch.log 2>&1
cat /root/repo/gpurun_out/zipbatch.log
