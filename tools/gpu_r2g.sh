#!/bin/bash
# pass g: trace_kernel queue-depth variants; the new bench.py (both arms); toggles table; staged-kernel ncu; GPU suite
set -u
mkdir -p gpurun_out
TAG=${1:-r2g}
for so in project3-cuda-path-tracer_b200/libptb200.so variants/*.so; do
  PTB_SO=$so PTB_B=16 timeout 120 python tools/gpu_time_synth.py 2>&1 | grep -E "TIMING|Error|error" | tail -2
done | tee gpurun_out/${TAG}_variants.txt
( time python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 6000 gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_reference.json
( time timeout 900 python tools/gpu_toggles.py ) > gpurun_out/${TAG}_toggles.txt 2>&1; tail -60 gpurun_out/${TAG}_toggles.txt
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log
