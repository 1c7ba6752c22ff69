#!/bin/bash
# round 2, call X: ncu captures of the one-row-per-lane Gauss-Jordan wrench kernels (1 warp at N = 11, 3 warps at N = 32)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
SRI_WRENCH_IMPL=multi ncu --set full --clock-control none --import-source on -k regex:wrench_local_solve_gj_multi -s 1 -c 1 -o gpurun_out/r2x_prof_gjm1 python tools/time_wrench.py 100000 11 > gpurun_out/r2x_ncu1.log 2>&1; tail -2 gpurun_out/r2x_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:wrench_local_solve_gj_multi -s 1 -c 1 -o gpurun_out/r2x_prof_gjm3 python tools/time_wrench.py 20000 32 > gpurun_out/r2x_ncu3.log 2>&1; tail -2 gpurun_out/r2x_ncu3.log
echo DONE
