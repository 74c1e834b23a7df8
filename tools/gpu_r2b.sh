#!/bin/bash
# Round-2 pass b: the multiplexed trace_kernel -- variants A/B on the synthetic 100 k-object scene, the whole GPU suite, sanitizer.
set -u
mkdir -p gpurun_out
TAG=${1:-r2b}
for so in project3-cuda-path-tracer_b200/libptb200.so variants/*.so; do
  PTB_SO=$so PTB_B=16 timeout 120 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2
done | tee gpurun_out/${TAG}_variants.txt
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -15 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 bash tools/gpu_sanitize.sh ${TAG} 2>&1 | tail -8
