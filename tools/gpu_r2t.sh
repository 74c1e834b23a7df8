#!/bin/bash
# pass t: 4-wide collapse of the BVH (-DPTB_TRACE_WIDE) against the shipped binary traversal; BVH tests with a wide build; ncu of one
set -u
mkdir -p gpurun_out
TAG=${1:-r2t}
for so in project3-cuda-path-tracer_b200/libptb200.so variants/*.so; do
  PTB_SO=$so PTB_B=16 timeout 120 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2
done | tee gpurun_out/${TAG}_variants.txt
W=${2:-variants/wide_s2.so}
( time PTB200_LIB=$PWD/$W timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_meshes.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest_wide.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_wide.log
PTB_SO=$W PTB_B=4 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 16 -c 3 -f -o gpurun_out/${TAG}_trace \
      python tools/gpu_time_synth.py > gpurun_out/${TAG}_trace_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_trace_ncu.log
