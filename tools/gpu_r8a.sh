#!/bin/bash
# pass 8a: the queue between phase 1 and the walk (v1, bounceSegmentQ2) against the queue between walk and shading (v0, -DPTB_QUEUE_CANDIDATES=0)
set -u
mkdir -p gpurun_out
TAG=${1:-r8a}
P=project3-cuda-path-tracer_b200
for b in 128 64; do for v in v0 v1 v0 v1; do PTB_B=$b PTB_SO=$P/libptb200_$v.so python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING; done; done | tee gpurun_out/${TAG}_candq_ab.txt
for v in v0 v1; do PTB_HQ=1 PTB_B=128 PTB_SO=$P/libptb200_$v.so python tools/gpu_time.py glass.txt glossy.txt 2>&1 | grep TIMING; PTB_HQ=1 PTB_B=16 PTB_SO=$P/libptb200_$v.so python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING; done | tee -a gpurun_out/${TAG}_candq_ab.txt
PTB200_LIB=$PWD/$P/libptb200_v1.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extensions.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -4 | tee -a gpurun_out/${TAG}_candq_ab.txt
