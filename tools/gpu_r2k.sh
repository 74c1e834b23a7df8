#!/bin/bash
# pass k: compute-sanitizer over the final kernels (trace_kernel, bounce0_shared_kernel, ordered deposit, staged + sort), headless toggles
set -u
mkdir -p gpurun_out/san
TAG=${1:-r2k}
python - <<'PY'
import sys; sys.path.insert(0, 'tests'); import scenes
open('gpurun_out/san/cornell_small.txt', 'w').write(scenes.cornell_text(res=(200, 200)))
open('gpurun_out/san/synth_small.txt', 'w').write(scenes.synthetic_text(2000, res=(160, 90), depth=8))
PY
H=project3-cuda-path-tracer_b200/ptb200_headless
for tool in memcheck racecheck synccheck; do
  for cfg in "cornell_small --batch 3 --iters 7" "synth_small --batch 2 --iters 4" "cornell_small --staged --sort --batch 2 --iters 4" "synth_small --staged --batch 2 --iters 2" "cornell_small --per-iteration --lazy-image --iters 3" "synth_small --cache-first-bounce --batch 2 --iters 4"; do
    set -- $cfg; s=$1; shift
    timeout -s KILL 300 compute-sanitizer --tool $tool --error-exitcode 7 $H gpurun_out/san/$s.txt "$@" --quiet --out gpurun_out/san/out_$tool > gpurun_out/san/${TAG}_$tool.log 2>&1
    echo "$tool [$cfg] rc=$? $(grep -E 'ERROR SUMMARY' gpurun_out/san/${TAG}_$tool.log | tail -1)"
  done
done | tee gpurun_out/${TAG}_sanitizer.txt
