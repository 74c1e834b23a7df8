#!/bin/bash
# pass w: shading side of BVH scenes (loadGeom fetches 64 instead of 160 bytes per shaded path): timing against HEAD, BVH / mesh / parity tests, ncu of bounce_kernel
set -u
mkdir -p gpurun_out
TAG=${1:-r2w}
for so in project3-cuda-path-tracer_b200/libptb200.so variants/*.so; do
  PTB_SO=$so PTB_B=16 timeout 120 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2
done | tee gpurun_out/${TAG}_variants.txt
PTB_B=0 timeout 200 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2 | tee -a gpurun_out/${TAG}_variants.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -4 gpurun_out/${TAG}_pytest_gpu.log
PTB_B=4 ncu --set full --clock-control none --import-source on -k regex:'bounce_kernel|bounce0_shared' -s 16 -c 3 -f -o gpurun_out/${TAG}_shade \
      python tools/gpu_time_synth.py > gpurun_out/${TAG}_shade_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_shade_ncu.log
