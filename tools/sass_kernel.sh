#!/bin/bash
# SASS of one kernel of a library, with source-line annotations:  tools/sass_kernel.sh LIB.so MANGLED_NAME > out.txt
T=$(mktemp -d); cd $T
cuobjdump -xelf all "$1" > /dev/null
nvdisasm -c -g ptb200.sm_100a.cubin 2>/dev/null | awk -v k="$2:" '$0==k{f=1} f{print} /^\/\/-----/{if(f&&n++>0)exit}'
