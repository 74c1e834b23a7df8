#!/bin/bash
# BASELINE.json configs[3] / configs[4] on N GPUs of one box (tools/gpu_multi_configs.py), largest N first with the
# reduced frame checked against a one-GPU render of the same iterations.   tools/gpu_multi.sh <tag> 8 4 2
set -u
mkdir -p gpurun_out
TAG=${1:-r01}; shift
FIRST=1
for N in "$@"; do
  CHECK=""; [ $FIRST = 1 ] && CHECK="--check"; FIRST=0
  timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 \
      tools/gpu_multi_configs.py $CHECK > gpurun_out/${TAG}_multi_n$N.jsonl 2> gpurun_out/${TAG}_multi_n$N.err
  echo "N=$N rc=$?"; cat gpurun_out/${TAG}_multi_n$N.jsonl; tail -3 gpurun_out/${TAG}_multi_n$N.err
done
