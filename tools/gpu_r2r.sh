#!/bin/bash
# pass r: compact pruning records + 256-bit primitive loads (A/B), default wavefront size, BVH tests
set -u
mkdir -p gpurun_out
TAG=${1:-r2r}
for so in project3-cuda-path-tracer_b200/libptb200.so variants/*.so; do
  PTB_SO=$so PTB_B=16 timeout 120 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2
done | tee gpurun_out/${TAG}_variants.txt
PTB_B=0 timeout 200 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2 | tee -a gpurun_out/${TAG}_variants.txt
( time timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_meshes.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest_bvh.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_bvh.log
