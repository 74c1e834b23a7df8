#!/bin/bash
# pass 5d: trace_kernel<SRC_SEGMENTS> with dynamically claimed path segments (v6, -DPTB_TRACE_DYNAMIC=1) against the static assignment (v0),
# segments of 1/4, 1/16, 1/32 of a trace CTA's share (PTB_GRID_MULT 4, 16, 32) on the 100k-object scene; BVH tests on the variant
set -u
mkdir -p gpurun_out
TAG=${1:-r5d}
P=project3-cuda-path-tracer_b200
for m in 4 16 32; do for v in v0 v6; do echo -n "mult $m: "; PTB_GRID_MULT=$m PTB_B=0 PTB_SO=$P/libptb200_$v.so timeout 300 python tools/gpu_time_synth.py 2>&1 | grep TIMING; done; done | tee gpurun_out/${TAG}_dynamic_ab.txt
for m in 8 16; do echo -n "mult $m: "; PTB_GRID_MULT=$m PTB_B=0 PTB_SO=$P/libptb200_v6.so timeout 300 python tools/gpu_time_synth.py 2>&1 | grep TIMING; done | tee -a gpurun_out/${TAG}_dynamic_ab.txt
PTB_GRID_MULT=16 PTB200_LIB=$PWD/$P/libptb200_v6.so timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_meshes.py -m gpu -q -x 2>&1 | tail -4 | tee -a gpurun_out/${TAG}_dynamic_ab.txt
