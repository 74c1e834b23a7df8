#!/bin/bash
# Round-2 first GPU pass: parity suite on HEAD, configs[4] on the shipped (far-pad) boxes against PTB_BVH_FAR_PAD=0,
# launch list + full ncu capture of the BVH kernels (the "before" of the trace_bvh_kernel redesign), path diff between the pads.
set -u
mkdir -p gpurun_out
TAG=${1:-r2a}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -3 gpurun_out/${TAG}_pytest_gpu.log
( PTB_B=16 python tools/gpu_time_synth.py; PTB_SO=project3-cuda-path-tracer_b200/libptb200_tuning.so PTB_BVH_FAR_PAD=0 PTB_B=16 python tools/gpu_time_synth.py ) > gpurun_out/${TAG}_synth_pad.txt 2>&1
cat gpurun_out/${TAG}_synth_pad.txt
PTB_B=4 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_synth_launches.csv \
    python tools/gpu_time_synth.py > gpurun_out/${TAG}_synth_under_ncu.log 2>&1
PTB_B=4 ncu --set full --clock-control none --import-source on -k regex:trace_bvh -s 17 -c 2 -f -o gpurun_out/${TAG}_trace \
    python tools/gpu_time_synth.py > gpurun_out/${TAG}_ncu_trace.log 2>&1
PTB_B=4 ncu --set full --clock-control none --import-source on -k regex:bounce_kernel -s 17 -c 2 -f -o gpurun_out/${TAG}_bvhbounce \
    python tools/gpu_time_synth.py > gpurun_out/${TAG}_ncu_bvhbounce.log 2>&1
( OLD_ALL_AXES=1 OLD_FAR_PAD=0 NEW_FAR_PAD=1 timeout 600 python tools/gpu_bvh_pathdiff.py 1 2 ) > gpurun_out/${TAG}_pathdiff.txt 2>&1
tail -3 gpurun_out/${TAG}_pathdiff.txt
ls -la gpurun_out | tail -12
