#!/bin/bash
# pass l: trace_pool_kernel (64 ray slots per warp in shared memory) -- timing, BVH tests with the pool build, ncu
set -u
mkdir -p gpurun_out
TAG=${1:-r2l}
for so in project3-cuda-path-tracer_b200/libptb200.so variants/*.so; do
  PTB_SO=$so PTB_B=16 timeout 120 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2
done | tee gpurun_out/${TAG}_variants.txt
( time PTB200_LIB=$PWD/variants/pool.so timeout 1200 python -m pytest tests/test_gpu_scale.py tests/test_gpu_meshes.py tests/test_gpu_parity.py tests/test_gpu_extensions.py -m gpu -q -x ) > gpurun_out/${TAG}_pytest_pool.log 2>&1
tail -6 gpurun_out/${TAG}_pytest_pool.log
PTB_SO=variants/pool.so PTB_B=4 ncu --set full --clock-control none --import-source on -k regex:trace_pool -s 16 -c 2 -f -o gpurun_out/${TAG}_pool \
      python tools/gpu_time_synth.py > gpurun_out/${TAG}_pool_ncu.log 2>&1
