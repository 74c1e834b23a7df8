#!/bin/bash
# pass 5c: bounce0_shared_kernel deals its tiles in units of one tile per warp (v5) against the release build: cornell at 128 / 64 / 16 / 1
# iterations per wavefront, sphere.txt, the 100k-object scene; parity tests on the variant
set -u
mkdir -p gpurun_out
TAG=${1:-r5c}
P=project3-cuda-path-tracer_b200
for b in 128 64 16 1; do for v in libptb200 libptb200_v5 libptb200 libptb200_v5; do PTB_B=$b PTB_SO=$P/$v.so python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING; done; done | tee gpurun_out/${TAG}_units_ab.txt
for v in libptb200 libptb200_v5; do PTB_B=128 PTB_SO=$P/$v.so python tools/gpu_time.py sphere.txt 2>&1 | grep TIMING; PTB_B=0 PTB_SO=$P/$v.so python tools/gpu_time_synth.py 2>&1 | grep TIMING; done | tee -a gpurun_out/${TAG}_units_ab.txt
PTB200_LIB=$PWD/$P/libptb200_v5.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_extensions.py tests/test_gpu_dropin.py -m gpu -q -x 2>&1 | tail -4 | tee -a gpurun_out/${TAG}_units_ab.txt
