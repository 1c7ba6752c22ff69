#!/bin/bash
# round 2, call R: ncu capture of the rolled Gauss-Jordan wrench kernel
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:wrench_local_solve_gj -s 1 -c 1 -o gpurun_out/r2r_prof_wrench_gj python tools/time_wrench.py 100000 16 > gpurun_out/r2r_ncu.log 2>&1; tail -2 gpurun_out/r2r_ncu.log
echo DONE
