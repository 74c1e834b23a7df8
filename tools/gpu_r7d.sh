#!/bin/bash
# pass 7d: PtbOptions.hitQueue in the release build: GPU suite, then off / on / auto on cornell (128, 64, 16 iterations per wavefront), glass, glossy
set -u
mkdir -p gpurun_out
TAG=${1:-r7d}
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${TAG}_pytest_gpu.log
for b in 128 64 16; do for h in 0 1 auto; do if [ $h = auto ]; then PTB_B=$b python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING; else PTB_HQ=$h PTB_B=$b python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING; fi; done; done | tee gpurun_out/${TAG}_queue_ab.txt
for h in 0 1; do PTB_HQ=$h PTB_B=128 python tools/gpu_time.py glass.txt glossy.txt 2>&1 | grep TIMING; done | tee -a gpurun_out/${TAG}_queue_ab.txt
PTB_B=128 python tools/gpu_time.py glass.txt glossy.txt 2>&1 | grep TIMING | tee -a gpurun_out/${TAG}_queue_ab.txt
