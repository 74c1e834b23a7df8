#!/bin/bash
# pass 4c: closestHitBrute<.., ONE> (scenes of <= 32 primitives: no chunk loop) A/B against HEAD on cornell at the bench's wavefront size,
# parity subset on the variant, sanitizer passes over the leave-test pipeline (release build)
set -u
mkdir -p gpurun_out
TAG=${1:-r4c}
P=project3-cuda-path-tracer_b200
for v in v0 v1 v0 v1; do PTB_B=128 PTB_SO=$P/libptb200_$v.so python tools/gpu_time.py cornell.txt sphere.txt 2>&1 | grep TIMING; done | tee gpurun_out/${TAG}_one_chunk_ab.txt
PTB200_LIB=$PWD/$P/libptb200_v1.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "leave_test or bench_configuration or pipelines_agree or bounce_loop_launch or batched" 2>&1 | tail -4 | tee -a gpurun_out/${TAG}_one_chunk_ab.txt
bash tools/gpu_sanitize.sh ${TAG}
