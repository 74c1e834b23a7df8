#!/bin/bash
# pass u, v: node format A/B; exact-test passes of one primitive type; whole GPU suite; ncu
set -u
mkdir -p gpurun_out
TAG=${1:-r2u}
for so in project3-cuda-path-tracer_b200/libptb200.so variants/*.so; do
  PTB_SO=$so PTB_B=16 timeout 120 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2
done | tee gpurun_out/${TAG}_variants.txt
PTB_B=0 timeout 200 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2 | tee -a gpurun_out/${TAG}_variants.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -6 gpurun_out/${TAG}_pytest_gpu.log
PTB_B=4 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 16 -c 3 -f -o gpurun_out/${TAG}_trace \
      python tools/gpu_time_synth.py > gpurun_out/${TAG}_trace_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_trace_ncu.log
