#!/bin/bash
# pass 4a: PtbOptions.leaveTest (closestHitBrute<LEAVE>): parity, A/B timing on cornell at the bench's wavefront size
set -u
mkdir -p gpurun_out
TAG=${1:-r4a}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "leave_test or bench_configuration or pipelines_agree or bounce_loop_launch" 2>&1 | tail -6
for L in 0 1 0 1; do PTB_B=128 PTB_LEAVE=$L python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING; done | tee gpurun_out/${TAG}_leave_ab.txt
