#!/bin/bash
# usage: tools/ncu_variant.sh <lib.so> <tag>   -- full ncu capture of the fused kernel for one build
LIB=$1; TAG=$2
export SRI_LIB_PATH=$LIB
CMD="python tools/variant_time.py 200000"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused16 -s 4 -c 1 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log
