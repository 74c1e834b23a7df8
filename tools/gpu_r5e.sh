#!/bin/bash
# pass 5e: trace_kernel<SRC_SEGMENTS> with per-warp dynamically claimed pieces of the path segments (v7: 512 rays, v8: 256, v9: 2048) against
# the static assignment (v0) on the 100k-object scene at 64 / 16 / 4 iterations per wavefront; BVH / mesh / multi tests on the variant
set -u
mkdir -p gpurun_out
TAG=${1:-r5e}
P=project3-cuda-path-tracer_b200
for v in v0 v7 v8 v9 v0 v7; do PTB_B=0 PTB_SO=$P/libptb200_$v.so timeout 300 python tools/gpu_time_synth.py 2>&1 | grep TIMING; done | tee gpurun_out/${TAG}_pieces_ab.txt
for b in 16 4; do for v in v0 v7; do PTB_B=$b PTB_SO=$P/libptb200_$v.so timeout 300 python tools/gpu_time_synth.py 2>&1 | grep TIMING; done; done | tee -a gpurun_out/${TAG}_pieces_ab.txt
PTB200_LIB=$PWD/$P/libptb200_v7.so timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_meshes.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -4 | tee -a gpurun_out/${TAG}_pieces_ab.txt
