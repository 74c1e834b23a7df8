#!/bin/bash
# ncu of the multiplexed trace_kernel (bounce-1 launch of the synthetic scene, B = 4): R = 2 (base) and R = 1
set -u
mkdir -p gpurun_out
TAG=${1:-r2c}
for v in base r1mb3; do
  PTB_SO=variants/$v.so PTB_B=4 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 17 -c 1 -f -o gpurun_out/${TAG}_${v} \
      python tools/gpu_time_synth.py > gpurun_out/${TAG}_${v}_ncu.log 2>&1
  tail -2 gpurun_out/${TAG}_${v}_ncu.log
done
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -8 gpurun_out/${TAG}_pytest_gpu.log
