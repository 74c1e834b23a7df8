#!/bin/bash
# tools/sass_fn.sh <lib.so> <mangled-name> : SASS of one kernel (instruction lines only, encodings stripped)
cuobjdump -sass "$1" | awk -v f="$2" '/Function :/{on=($3==f)} on' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's/^\s*//; s#/\* 0x[0-9a-f]* \*/##; s/\s*$//'
