#!/bin/bash
# A/B of the one-launch-for-bounces-1.. kernel (PTB_LOOP_BOUNCES) and the CTA count multiplier: gpurun_out/TAG_loop_ab.txt
set -u
mkdir -p gpurun_out
TAG=$1
for cfg in "0 4" "1 4" "1 2" "1 8" "1 16" "0 4" "1 4"; do
  set -- $cfg
  PTB_LOOP_BOUNCES=$1 PTB_GRID_MULT=$2 PTB_B=64 timeout 300 python tools/gpu_time.py cornell.txt sphere.txt 2>&1 | grep TIMING | sed "s/^default/loop=$1/"
done | tee gpurun_out/${TAG}_loop_ab.txt
