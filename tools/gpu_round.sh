#!/bin/bash
# One GPU-box pass: parity tests, bench (both arms), the skeleton CUDA baseline, ncu launch list and
# one full ncu capture of the bounce kernels.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${TAG}_smi.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -5 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 3000 gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_reference.json
python -c "
import sys; sys.path.insert(0,'tests'); import scenes; scenes.write_all('gpurun_out/scenes')"
if [ -x oracle/_ref/skeleton_bench ]; then
  oracle/_ref/skeleton_bench gpurun_out/scenes/cornell.txt 1000 50 > gpurun_out/${TAG}_skeleton_cornell.json 2>&1; cat gpurun_out/${TAG}_skeleton_cornell.json
  oracle/_ref/skeleton_bench gpurun_out/scenes/sphere.txt 1000 50 > gpurun_out/${TAG}_skeleton_sphere.json 2>&1; cat gpurun_out/${TAG}_skeleton_sphere.json
fi
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --spp-per-step 16 --no-cpu-baseline > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
# one full capture of the two bounce launches (bounce 0, bounces 1..) of the timed step's 128-iteration wavefront (bench.py BATCH_ITERS)
ncu --set full --clock-control none --import-source on -k regex:bounce -s 6 -c 2 -f -o gpurun_out/${TAG}_bounce \
    python bench.py --steps 1 --warmup 3 --spp-per-step 128 --no-cpu-baseline > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out
