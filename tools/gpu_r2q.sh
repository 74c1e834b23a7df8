#!/bin/bash
# pass q: L1 policies of the node / primitive loads, wavefront size, per-launch durations of one wavefront of the 100k scene
set -u
mkdir -p gpurun_out
TAG=${1:-r2q}
for so in project3-cuda-path-tracer_b200/libptb200.so variants/*.so; do
  PTB_SO=$so PTB_B=16 timeout 120 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2
done | tee gpurun_out/${TAG}_variants.txt
for b in 8 32 64; do PTB_B=$b timeout 200 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2; done | tee -a gpurun_out/${TAG}_variants.txt
PTB_B=16 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_synth_launches.csv python tools/gpu_time_synth.py > gpurun_out/${TAG}_synth_under_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_synth_under_ncu.log
