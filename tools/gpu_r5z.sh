#!/bin/bash
# pass 5z: whole-repo verification on HEAD (one-chunk brute-force search): tools/gpu_r2s.sh, then the wavefront size of the 100k-object scene
set -u
TAG=${1:-r5z}
bash tools/gpu_r2s.sh $TAG
for b in 0 128 0 128; do PTB_B=$b python tools/gpu_time_synth.py 2>&1 | grep TIMING; done | tee gpurun_out/${TAG}_synth_wavefront.txt
