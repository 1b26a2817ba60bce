#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "zip_batch" > gpurun_out/pytest_gpu_x.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_x.log; tail -6 gpurun_out/pytest_gpu_x.log
mkdir -p /dev/shm/zb && cd /dev/shm/zb
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda c4.zip 100000 65536 6 cuda > /root/repo/gpurun_out/zipx.log 2>&1
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda c4.zip 100000 65536 6 extract >> /root/repo/gpurun_out/zipx.log 2>&1
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda c4.zip 100000 65536 6 extract >> /root/repo/gpurun_out/zipx.log 2>&1
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda r.zip 6000 65536 6 ref >> /root/repo/gpurun_out/zipx.log 2>&1
timeout 600 /root/repo/oracle/_ref/zipbatch_cuda r.zip 6000 65536 6 extract_ref >> /root/repo/gpurun_out/zipx.log 2>&1
cut -c1-420 /root/repo/gpurun_out/zipx.log
rm -rf /dev/shm/zb
