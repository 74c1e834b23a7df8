#!/bin/bash
# pass 5a: (1) phase-1 key FMA carries |d|, one predicate for reach && !skip (v3) against the release build; (2) grid of the fused
# pipeline against the warp imbalance of bounce0_shared_kernel (80 000 tiles over 4 736 CTAs x 8 warps = 2.1 tiles per warp)
set -u
mkdir -p gpurun_out
TAG=${1:-r5a}
P=project3-cuda-path-tracer_b200
for v in libptb200 libptb200_v3 libptb200 libptb200_v3; do PTB_B=128 PTB_SO=$P/$v.so python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING; done | tee gpurun_out/${TAG}_keys_ab.txt
for g in 4736 5000 3334 2500 10000 6667 4736 5000; do echo -n "grid $g: "; PTB_GRID_FUSED=$g PTB_B=128 PTB_SO=$P/libptb200_v3.so python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING; done | tee gpurun_out/${TAG}_grid_ab.txt
PTB200_LIB=$PWD/$P/libptb200_v3.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "leave_test or bench_configuration or pipelines_agree or one_chunk or batched" 2>&1 | tail -4 | tee -a gpurun_out/${TAG}_keys_ab.txt
