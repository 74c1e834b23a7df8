#!/bin/bash
# pass 9a: why is the hit queue slower on the glass scene?  ncu of one bounce_loop launch with the queue off and on
set -u
mkdir -p gpurun_out
TAG=${1:-r9a}
for h in 0 1; do
  PTB_HQ=$h PTB_B=128 ncu --set full --clock-control none --import-source on -k regex:bounce_loop -s 2 -c 1 -f -o gpurun_out/${TAG}_glass_hq$h \
      python tools/gpu_time.py glass.txt > gpurun_out/${TAG}_glass_hq$h.log 2>&1
done
ls -la gpurun_out | grep ${TAG}
