#!/bin/bash
# One GPU call: tests, bench, ncu launch list and one full capture of the fused kernel.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3500 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --rods 200000"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:fused16_dmma -s 3 -c 1 -o gpurun_out/prof_fused16_dmma $CMD > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out | tail -8
