#!/bin/bash
# quick parity/timing check + one full ncu capture of the 8 bounce kernels of one cornell wavefront
TAG=${1:-dev}
mkdir -p gpurun_out
timeout 500 python tools/gpu_quick.py 2>&1 | tee gpurun_out/${TAG}_quick.log | grep -E "TIMING|False|alive"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bounce_kernel -s 8 -c 8 -f -o gpurun_out/${TAG}_bounce \
    python tools/gpu_one_batch.py 2 > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
