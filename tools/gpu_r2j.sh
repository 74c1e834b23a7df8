#!/bin/bash
# pass j: bounce 0 shared across the slots of a wavefront + two-level sort scatter: GPU suite, bench (both arms), toggles, ncu summaries
set -u
mkdir -p gpurun_out
TAG=${1:-r2j}
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -6 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
( time python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
( time timeout 900 python tools/gpu_toggles.py ) > gpurun_out/${TAG}_toggles.txt 2>&1; head -12 gpurun_out/${TAG}_toggles.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --spp-per-step 128 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'bounce0_shared|bounce_loop' -s 6 -c 2 -f -o gpurun_out/${TAG}_bounce \
    python bench.py --steps 1 --warmup 3 --spp-per-step 128 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_full.log 2>&1
python -c "
import sys; sys.path.insert(0,'tests'); import scenes; scenes.write_all('gpurun_out/scenes')"
H=project3-cuda-path-tracer_b200/ptb200_headless
ncu --set full --clock-control none --import-source on -k regex:'raygen_kernel|sort_scatter|deposit_reduce' -c 4 -f -o gpurun_out/${TAG}_staged2 \
      $H gpurun_out/scenes/cornell.txt --staged --sort --batch 8 --iters 8 --quiet --out gpurun_out/${TAG}_staged_img > gpurun_out/${TAG}_staged2_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'gather_kernel' -c 1 -f -o gpurun_out/${TAG}_gather \
      $H gpurun_out/scenes/cornell.txt --staged --no-compaction --atomic-deposit --batch 8 --iters 8 --quiet --out gpurun_out/${TAG}_gather_img > gpurun_out/${TAG}_gather_ncu.log 2>&1
ls gpurun_out | grep ${TAG} | tr '\n' ' '
