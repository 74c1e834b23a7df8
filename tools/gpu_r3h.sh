#!/bin/bash
# pass 3h: the reference's per-iteration call through the C++ shim (ptb200_headless --per-iteration: pageable std::vector image, pinned by the
# library), with and without the look-ahead, strict and lazy; sanitizer over the look-ahead path
set -u
mkdir -p gpurun_out
TAG=${1:-r3h}
python -c "
import sys; sys.path.insert(0,'tests'); import scenes
open('gpurun_out/cornell_2000.txt','w').write(scenes.cornell_text(iterations=2000))
open('gpurun_out/cornell_small.txt','w').write(scenes.cornell_text(res=(200,200), iterations=40))"
H=project3-cuda-path-tracer_b200/ptb200_headless
for opts in "--no-lookahead" "--no-lookahead --lazy-image" "" "--lazy-image"; do
  echo "== --per-iteration $opts"; $H gpurun_out/cornell_2000.txt --per-iteration $opts --quiet --out gpurun_out/${TAG}_img 2>&1 | tail -1
  $H gpurun_out/cornell_2000.txt --per-iteration $opts --out gpurun_out/${TAG}_img 2>&1 | grep "iterations in"
done | tee gpurun_out/${TAG}_shim.txt
for tool in memcheck racecheck synccheck; do
  timeout -s KILL 300 compute-sanitizer --tool $tool --error-exitcode 7 $H gpurun_out/cornell_small.txt --per-iteration --batch 8 --quiet --out gpurun_out/${TAG}_san > gpurun_out/${TAG}_$tool.log 2>&1
  echo "$tool [--per-iteration --batch 8 (look-ahead)] rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/${TAG}_$tool.log | tail -1)"
done | tee -a gpurun_out/${TAG}_shim.txt
