#!/bin/bash
# one full ncu capture of the first BVH trace kernels (synthetic 100k-object scene)
TAG=${1:-dev}
mkdir -p gpurun_out
timeout 800 ncu --set full --clock-control none --import-source on -k regex:trace_bvh_kernel -s 16 -c 3 -f -o gpurun_out/${TAG}_bvh \
    python tools/gpu_bvh_batch.py 2 2 > gpurun_out/${TAG}_bvh_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_bvh_ncu.log
