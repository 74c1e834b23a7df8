#!/bin/bash
# pass n: trace_kernel with the pointer stack (sentinel-terminated, block spills) -- variants, BVH tests (also with an 8-entry stack
# so that the spill path runs), whole GPU suite, ncu of the bounce-1 launch
set -u
mkdir -p gpurun_out
TAG=${1:-r2n}
for so in project3-cuda-path-tracer_b200/libptb200.so variants/*.so; do
  PTB_SO=$so PTB_B=16 timeout 120 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2
done | tee gpurun_out/${TAG}_variants.txt
( time PTB200_LIB=$PWD/variants/k8.so timeout 900 python -m pytest tests/test_gpu_scale.py tests/test_gpu_meshes.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest_k8.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_k8.log
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -6 gpurun_out/${TAG}_pytest_gpu.log
PTB_B=4 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 16 -c 3 -f -o gpurun_out/${TAG}_trace \
      python tools/gpu_time_synth.py > gpurun_out/${TAG}_trace_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_trace_ncu.log
