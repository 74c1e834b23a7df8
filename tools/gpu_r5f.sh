#!/bin/bash
# pass 5f: piece size adapted to the launch (32 .. max rays, ~8 pieces per warp): max 256 (v7), 128 (v8), 64 (v9) against CTA-owned segments (v0),
# 100k-object scene at 64 / 16 / 4 / 1 iterations per wavefront; BVH / mesh / parity tests on v7
set -u
mkdir -p gpurun_out
TAG=${1:-r5f}
P=project3-cuda-path-tracer_b200
for b in 0 16 4 1; do for v in v0 v7 v8 v9; do PTB_B=$b PTB_SO=$P/libptb200_$v.so timeout 300 python tools/gpu_time_synth.py 2>&1 | grep TIMING; done; done | tee gpurun_out/${TAG}_pieces_ab.txt
PTB200_LIB=$PWD/$P/libptb200_v7.so timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_meshes.py -m gpu -q -x 2>&1 | tail -4 | tee -a gpurun_out/${TAG}_pieces_ab.txt
