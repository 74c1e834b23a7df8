#!/bin/bash
# pass 9d: compute-sanitizer racecheck over the two new shared-memory protocols: the hit queue of bounce_loop_kernel (cornell, --batch 64 so that
# AUTO switches it on at 200 x 200) and the claimed pieces of trace_kernel (2 000-object BVH scene)
set -u
mkdir -p gpurun_out/san
python - <<'PY'
import sys; sys.path.insert(0, 'tests'); import scenes
open('gpurun_out/san/cornell_small.txt', 'w').write(scenes.cornell_text(res=(200, 200)))
open('gpurun_out/san/synth_small.txt', 'w').write(scenes.synthetic_text(2000, res=(160, 90), depth=8))
PY
H=project3-cuda-path-tracer_b200/ptb200_headless
( timeout -s KILL 40 compute-sanitizer --tool racecheck --error-exitcode 7 $H gpurun_out/san/synth_small.txt --iters 4 --batch 2 --quiet --out gpurun_out/san/s > gpurun_out/san/r9d_synth_racecheck.log 2>&1; echo "racecheck synth_small rc=$? $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/san/r9d_synth_racecheck.log | tail -1)" ) &
( timeout -s KILL 40 compute-sanitizer --tool racecheck --error-exitcode 7 $H gpurun_out/san/cornell_small.txt --iters 64 --batch 64 --quiet --out gpurun_out/san/c > gpurun_out/san/r9d_cornell_racecheck.log 2>&1; echo "racecheck cornell_small batch 64 rc=$? $(grep -E 'RACECHECK SUMMARY|ERROR SUMMARY' gpurun_out/san/r9d_cornell_racecheck.log | tail -1)" ) &
wait
