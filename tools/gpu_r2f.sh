#!/bin/bash
# pass f: speculative trace_kernel (one ray per lane, leaf queue, pop-time culling) -- variants, ncu, the whole GPU suite
set -u
mkdir -p gpurun_out
TAG=${1:-r2f}
for so in project3-cuda-path-tracer_b200/libptb200.so variants/*.so; do
  PTB_SO=$so PTB_B=16 timeout 120 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2
done | tee gpurun_out/${TAG}_variants.txt
PTB_B=4 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 17 -c 1 -f -o gpurun_out/${TAG}_trace \
      python tools/gpu_time_synth.py > gpurun_out/${TAG}_trace_ncu.log 2>&1
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log
