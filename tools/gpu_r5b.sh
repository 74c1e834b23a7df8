#!/bin/bash
# pass 5b: grid sweep of the fused pipeline on cornell (128 and 64 iterations per wavefront) and on the 4K cornell (16 per wavefront)
set -u
mkdir -p gpurun_out
TAG=${1:-r5b}
P=project3-cuda-path-tracer_b200
for g in 4736 8000 9472 10000 11840 13334 14208 16000 18944 20000 10000 4736; do echo -n "grid $g: "; PTB_GRID_FUSED=$g PTB_B=128 PTB_SO=$P/libptb200_v3.so python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING; done | tee gpurun_out/${TAG}_grid_ab.txt
for g in 4736 5000 9472 10000; do echo -n "B=64 grid $g: "; PTB_GRID_FUSED=$g PTB_B=64 PTB_SO=$P/libptb200_v3.so python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING; done | tee -a gpurun_out/${TAG}_grid_ab.txt
