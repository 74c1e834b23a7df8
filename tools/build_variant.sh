#!/bin/bash
# builds a tuning variant of libptb200.so:  tools/build_variant.sh OUT.so [-DPTB_TRACE_RAYS=.. -DPTB_MINBLOCKS=.. ...]
OUT=$1; shift
C=${PTB_CSRC:-/root/repo/project3-cuda-path-tracer_b200/csrc}
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -I/root/repo/include -std=c++17 -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC,-O2,-ffp-contract=off -DPTB_TUNING "$@" -shared -o $OUT $C/ptb200.cu $C/ptb_bvh.cpp $C/host/scene.cpp $C/host/utilities.cpp $C/host/image.cpp $C/host/pathtrace.cpp $C/host/host_abi.cpp -lz -Xptxas -v 2>&1 \
  | grep -A2 -E "trace_kernelILi1ELb0ELb0|bounce_loop_kernelILb0" | grep -E "Used|spill" | tr '\n' ' '; echo " <- $OUT $@"
