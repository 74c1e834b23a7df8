#!/bin/bash
# A/B timing of tuning variants of libptb200.so (built with tools/build_variant.sh into variants/) on one GPU box:
#   tools/gpu_variants.sh TAG variants/a.so variants/b.so ...     -> gpurun_out/TAG_variants.txt
set -u
mkdir -p gpurun_out
TAG=$1; shift
for so in "$@"; do
  for rep in 1 2; do
    PTB_SO=$so PTB_B=64 timeout 300 python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING
  done
done | tee gpurun_out/${TAG}_variants.txt
