#!/bin/bash
# pass h (2 GPUs): the library's NCCL call sites -- ptb_multi_* test, headless --gpus, bench.py under torchrun
set -u
mkdir -p gpurun_out
TAG=${1:-r2h}
nvidia-smi -L
( time timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_dropin.py -m gpu -q ) > gpurun_out/${TAG}_pytest_multi.log 2>&1; tail -8 gpurun_out/${TAG}_pytest_multi.log
python -c "
import sys; sys.path.insert(0,'tests'); import scenes; scenes.write_all('gpurun_out/scenes')"
H=project3-cuda-path-tracer_b200/ptb200_headless
( $H gpurun_out/scenes/cornell.txt --iters 512 --out gpurun_out/${TAG}_g1 ; NCCL_DEBUG=INFO $H gpurun_out/scenes/cornell.txt --iters 512 --gpus 2 --out gpurun_out/${TAG}_g2 ) > gpurun_out/${TAG}_headless.log 2>&1
grep -E "iterations in|GPUs:|NCCL INFO.*(Connected|Channel 00|comm 0x.*rank)" gpurun_out/${TAG}_headless.log | head -12
cmp gpurun_out/${TAG}_g1.png gpurun_out/${TAG}_g2.png && echo "1-GPU and 2-GPU PNG identical"
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 ) > gpurun_out/${TAG}_bench_n2.json 2> gpurun_out/${TAG}_bench_n2.err
tail -c 2500 gpurun_out/${TAG}_bench_n2.json; tail -5 gpurun_out/${TAG}_bench_n2.err
