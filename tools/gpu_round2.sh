#!/bin/bash
# second evidence round: high-resolution kernels, Newton, stage kernels (+ ncu of the stage kernels)
mkdir -p gpurun_out
python tools/time_highres.py > gpurun_out/highres.jsonl 2>&1; cat gpurun_out/highres.jsonl
SRI_FUSED16_IMPL=scalar python tools/time_highres.py > gpurun_out/highres_scalar.jsonl 2>&1
python tools/bench_newton.py --rods 100000 > gpurun_out/newton.jsonl 2>&1; tail -1 gpurun_out/newton.jsonl | cut -c1-400
python tools/time_stages.py 16 2000000 > gpurun_out/stage_kernels.jsonl 2>&1; cat gpurun_out/stage_kernels.jsonl
ncu --set full --clock-control none -k regex:stage_dmma -s 6 -c 3 -o gpurun_out/prof_stages python tools/time_stages.py 16 400000 > gpurun_out/ncu_stages.log 2>&1
tail -n 2 gpurun_out/ncu_stages.log
