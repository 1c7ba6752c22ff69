#!/bin/bash
# round 2, run AB: static-order first pass of the local-frame statics: how many rods are handed back, and the time with the bound off / at 0
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
for N in 11 16 32; do
  B=200000; [ $N -gt 16 ] && B=40000
  for G in 4 1e300 0; do
    SRI_DMMA_GROWTH=$G timeout 120 python tools/time_wrench.py $B $N 2>/dev/null | head -1 | sed "s/^/growth=$G /" >> gpurun_out/r2ab_wrench.jsonl
  done
done
cut -c1-260 gpurun_out/r2ab_wrench.jsonl
