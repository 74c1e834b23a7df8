#!/bin/bash
# pass i: trace_kernel micro-variants; ncu of the default trace_kernel and of the staged pipeline's kernels; B = 1 grid sizing
set -u
mkdir -p gpurun_out
TAG=${1:-r2i}
for so in project3-cuda-path-tracer_b200/libptb200.so variants/*.so; do
  PTB_SO=$so PTB_B=16 timeout 120 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2
done | tee gpurun_out/${TAG}_variants.txt
for mult in 1 2 4 8; do
  PTB_SO=project3-cuda-path-tracer_b200/libptb200_tuning.so PTB_GRID_MULT=$mult PTB_B=1 timeout 120 python tools/gpu_time.py cornell.txt 2>&1 | grep TIMING
done | tee gpurun_out/${TAG}_b1_grid.txt
PTB_B=4 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 16 -c 3 -f -o gpurun_out/${TAG}_trace \
      python tools/gpu_time_synth.py > gpurun_out/${TAG}_trace_ncu.log 2>&1
python -c "
import sys; sys.path.insert(0,'tests'); import scenes; scenes.write_all('gpurun_out/scenes')"
H=project3-cuda-path-tracer_b200/ptb200_headless
ncu --set full --clock-control none --import-source on -k regex:'raygen_kernel|intersect_kernel|sort_|shade_kernel|compact_kernel|deposit_reduce' -s 9 -c 9 -f -o gpurun_out/${TAG}_staged \
      $H gpurun_out/scenes/cornell.txt --staged --sort --batch 8 --iters 16 --quiet --out gpurun_out/${TAG}_staged_img > gpurun_out/${TAG}_staged_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gather_kernel|carry_count' -c 2 -f -o gpurun_out/${TAG}_gather \
      $H gpurun_out/scenes/cornell.txt --staged --no-compaction --atomic-deposit --batch 8 --iters 8 --quiet --out gpurun_out/${TAG}_gather_img > gpurun_out/${TAG}_gather_ncu.log 2>&1
ls -la gpurun_out | grep ${TAG}
