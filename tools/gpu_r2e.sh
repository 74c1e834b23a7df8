#!/bin/bash
set -u
mkdir -p gpurun_out
TAG=${1:-r2e}
for v in base r1mb4; do
  PTB_SO=variants/$v.so PTB_B=4 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 17 -c 1 -f -o gpurun_out/${TAG}_${v} \
      python tools/gpu_time_synth.py > gpurun_out/${TAG}_${v}_ncu.log 2>&1
  tail -1 gpurun_out/${TAG}_${v}_ncu.log
done
