#!/bin/bash
# pass s: whole-repo verification on HEAD (pointer-stack trace_kernel, 256-bit node loads, bigger default wavefront):
# GPU suite, smoke, bench (both arms), cornell wavefront size A/B, launch list, ncu full of the bounce kernels and of trace_kernel
set -u
mkdir -p gpurun_out
TAG=${1:-r2s}
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/${TAG}_pytest_gpu.log 2>&1
tail -6 gpurun_out/${TAG}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
( time python bench.py ) > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -3 gpurun_out/${TAG}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
for b in 64 256; do python bench.py --no-extras --no-cpu-baseline --batch-iters $b 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch_iters', d['config'].get('batch_iters'), 'value %.4e'%d['value'], 'ms_per_step', d['ms_per_step'], 'e2e %.4e'%d['e2e']['value'])"; done | tee gpurun_out/${TAG}_cornell_wavefront.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --spp-per-step 128 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'bounce0_shared|bounce_loop' -s 6 -c 2 -f -o gpurun_out/${TAG}_bounce \
    python bench.py --steps 1 --warmup 3 --spp-per-step 128 --no-cpu-baseline --no-extras > gpurun_out/${TAG}_ncu_full.log 2>&1
PTB_B=4 ncu --set full --clock-control none --import-source on -k regex:trace_kernel -s 16 -c 3 -f -o gpurun_out/${TAG}_trace \
      python tools/gpu_time_synth.py > gpurun_out/${TAG}_trace_ncu.log 2>&1
ls gpurun_out | grep ${TAG} | tr '\n' ' '
