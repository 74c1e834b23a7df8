#!/bin/bash
# Scaling pass on one multi-GPU box: bench.py at N = 1, 2, ... (argument list), one rank per GPU over NCCL.
set -u
mkdir -p gpurun_out
TAG=${1:-r01}; shift
for N in "$@"; do
  if [ "$N" = "1" ]; then
    timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_scale_n1.json 2> gpurun_out/${TAG}_scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
        bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
  fi
  tail -c 1500 gpurun_out/${TAG}_scale_n$N.json; tail -3 gpurun_out/${TAG}_scale_n$N.err
done
