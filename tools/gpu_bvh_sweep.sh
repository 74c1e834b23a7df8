#!/bin/bash
# parity of the fused BVH path, then synthetic-scene timing for a few BVH leaf sizes
mkdir -p gpurun_out
for leaf in ${LEAVES:-4 2 1}; do
  echo "== PTB_BVH_LEAF=$leaf"
  PTB_BVH_LEAF=$leaf timeout 600 python tools/gpu_bvh_check.py 2>&1 | grep -E "bit for bit|TIMING"
done
