#!/bin/bash
# pass 7c: hit queue with the ray read again (v1) / carried through the queue (v2) against the chunk-at-a-time loop (v0)
set -u
mkdir -p gpurun_out
TAG=${1:-r7c}
P=project3-cuda-path-tracer_b200
for b in 128 16 1; do for v in v0 v1 v2; do PTB_B=$b PTB_SO=$P/libptb200_$v.so python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING; done; done | tee gpurun_out/${TAG}_queue_ab.txt
for v in v0 v1 v2; do PTB_B=128 PTB_SO=$P/libptb200_$v.so python tools/gpu_time.py glass.txt glossy.txt 2>&1 | grep TIMING; done | tee -a gpurun_out/${TAG}_queue_ab.txt
