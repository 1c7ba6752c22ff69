#!/bin/bash
# round 2, call AC: ncu captures of the static-order Gauss-Jordan wrench kernels (1 warp at N = 11, 2 warps at N = 16)
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:wrench_local_solve_gj_static -s 1 -c 1 -o gpurun_out/r2ac_prof_gjs1 python tools/time_wrench.py 100000 11 > gpurun_out/r2ac_ncu1.log 2>&1; tail -2 gpurun_out/r2ac_ncu1.log
ncu --set full --clock-control none --import-source on -k regex:wrench_local_solve_gj_static -s 1 -c 1 -o gpurun_out/r2ac_prof_gjs2 python tools/time_wrench.py 100000 16 > gpurun_out/r2ac_ncu2.log 2>&1; tail -2 gpurun_out/r2ac_ncu2.log
echo DONE
