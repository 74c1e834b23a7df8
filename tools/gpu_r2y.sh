#!/bin/bash
# pass y (N GPUs of one box): the library's NCCL path on HEAD -- ptb_multi_* test, headless --gpus N against 1 GPU, bench.py at 1, 2, .. N
# (weak scaling of the headline + the strong-scaling leg of configs[3]), configs[4] on N GPUs.   tools/gpu_r2y.sh <tag> <N>
set -u
mkdir -p gpurun_out
TAG=${1:-r2y}; NMAX=${2:-2}
nvidia-smi -L | tee gpurun_out/${TAG}_gpus.txt
( time timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/${TAG}_pytest_multi.log 2>&1; tail -4 gpurun_out/${TAG}_pytest_multi.log
python -c "
import sys; sys.path.insert(0,'tests'); import scenes; scenes.write_all('gpurun_out/scenes')"
H=project3-cuda-path-tracer_b200/ptb200_headless
( $H gpurun_out/scenes/cornell.txt --iters 1024 --out gpurun_out/${TAG}_g1 ; NCCL_DEBUG=INFO $H gpurun_out/scenes/cornell.txt --iters 1024 --gpus $NMAX --out gpurun_out/${TAG}_gN ) > gpurun_out/${TAG}_headless.log 2>&1
grep -E "iterations in|GPUs:|NCCL INFO.*(Connected all|comm 0x.*rank)" gpurun_out/${TAG}_headless.log | head -14
cmp gpurun_out/${TAG}_g1.png gpurun_out/${TAG}_gN.png && echo "1-GPU and $NMAX-GPU PNG identical"
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_scale_n1.json 2> gpurun_out/${TAG}_scale_n1.err
for N in 2 4 8; do
  [ $N -le $NMAX ] || continue
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
done
python - <<PY
import json, glob
base = None
for n in (1, 2, 4, 8):
    try: d = json.loads(open('gpurun_out/${TAG}_scale_n%d.json' % n).read().strip().splitlines()[-1])
    except Exception as e: continue
    if n == 1: base = d['value']
    ss = d.get('strong_scaling') or {}
    print('N=%d value %.4e seg/s  weak efficiency %.4f  e2e %.4e  strong 4K 5000spp %.1f ms (reduce %.2f ms)  comm ranks %s' % (n, d['value'], d['value'] / (n * base), d['e2e']['value'], ss.get('ms_per_frame', float('nan')), ss.get('reduce_ms_rank0', float('nan')), d.get('comm_nranks_seen')))
PY
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NMAX --master-addr 127.0.0.1 --master-port 29517 \
      tools/gpu_multi_configs.py --check synthetic > gpurun_out/${TAG}_multi_n$NMAX.jsonl 2> gpurun_out/${TAG}_multi_n$NMAX.err
cat gpurun_out/${TAG}_multi_n$NMAX.jsonl | cut -c1-700; tail -2 gpurun_out/${TAG}_multi_n$NMAX.err
