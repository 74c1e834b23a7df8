#!/bin/bash
# pass 6a: bounce0_shared_kernel shares a CTA's tiles out by rows (v1) against whole tiles per warp (v0): cornell at 128 / 64 / 16 / 1 iterations
# per wavefront, sphere.txt, 4K cornell, the 100k-object scene; parity / extension / drop-in tests on the variant
set -u
mkdir -p gpurun_out
TAG=${1:-r6a}
P=project3-cuda-path-tracer_b200
for b in 128 64 16 1; do for v in v0 v1 v0 v1; do PTB_B=$b PTB_SO=$P/libptb200_$v.so python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING; done; done | tee gpurun_out/${TAG}_rows_ab.txt
for v in v0 v1; do PTB_B=128 PTB_SO=$P/libptb200_$v.so python tools/gpu_time.py sphere.txt 2>&1 | grep TIMING; PTB_B=0 PTB_SO=$P/libptb200_$v.so python tools/gpu_time_synth.py 2>&1 | grep TIMING; done | tee -a gpurun_out/${TAG}_rows_ab.txt
PTB200_LIB=$PWD/$P/libptb200_v1.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extensions.py tests/test_gpu_dropin.py tests/test_gpu_scale.py -m gpu -q -x 2>&1 | tail -4 | tee -a gpurun_out/${TAG}_rows_ab.txt
