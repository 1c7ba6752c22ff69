#!/bin/bash
# round 2, call E: Jacobian (node-per-lane passes) tests + timing + ncu, write-only roof calibration
mkdir -p gpurun_out
python -m pytest tests/test_gpu_newton.py tests/test_gpu_boundary.py -m gpu -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2e_pytest.log
python tools/bench_jacobian.py | tee -a gpurun_out/r2e_jacobian.jsonl
SRI_LIB_PATH=$PWD/tools/_variants/libsri_jac4.so python tools/bench_jacobian.py | tee -a gpurun_out/r2e_jacobian.jsonl
python tools/bench_jacobian.py 12500 | tee -a gpurun_out/r2e_jacobian.jsonl
python tools/newton_once.py; python tools/newton_once.py 12500
python tools/time_stages.py 16 2000000 | grep -E "nofbar|memset" | tee -a gpurun_out/r2e_stage_noload.jsonl
ncu --set full --clock-control none --import-source on -k regex:shape_jacobian_dmma -s 2 -c 1 -o gpurun_out/r2e_prof_jacobian python tools/bench_jacobian.py > gpurun_out/r2e_ncu_jac.log 2>&1; tail -2 gpurun_out/r2e_ncu_jac.log
echo DONE
