#!/bin/bash
# compute-sanitizer passes over small renders through the headless driver: synccheck + memcheck + racecheck on the brute-force
# loop (cornell 200x200, 4 iterations, batch 2) and on the BVH pipeline (2 000-object synthetic scene 160x90).
set -u
mkdir -p gpurun_out/san
TAG=${1:-r02}
python - <<'PY'
import sys; sys.path.insert(0, 'tests'); import scenes
open('gpurun_out/san/cornell_small.txt', 'w').write(scenes.cornell_text(res=(200, 200)))
open('gpurun_out/san/synth_small.txt', 'w').write(scenes.synthetic_text(2000, res=(160, 90), depth=8))
PY
H=project3-cuda-path-tracer_b200/ptb200_headless
for tool in synccheck memcheck racecheck; do
  for s in cornell_small synth_small; do
    timeout -s KILL 300 compute-sanitizer --tool $tool --error-exitcode 7 $H gpurun_out/san/$s.txt --iters 4 --batch 2 --quiet --out gpurun_out/san/${s}_$tool > gpurun_out/san/${TAG}_${s}_$tool.log 2>&1
    echo "$tool $s rc=$? $(grep -E 'ERROR SUMMARY|Error' gpurun_out/san/${TAG}_${s}_$tool.log | tail -1)"
  done
done | tee gpurun_out/${TAG}_sanitizer.txt
