#!/bin/bash
# round 2, call F: Jacobian with cp.async prefetch (3 occupancy variants), no-load stress kernel v2
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2f_pytest.log
python tools/bench_jacobian.py | tee -a gpurun_out/r2f_jacobian.jsonl
for mb in 2 4; do SRI_LIB_PATH=$PWD/tools/_variants/libsri_jac$mb.so python tools/bench_jacobian.py | tee -a gpurun_out/r2f_jacobian.jsonl; done
python tools/bench_jacobian.py 12500 | tee -a gpurun_out/r2f_jacobian.jsonl
python tools/newton_once.py; python tools/newton_once.py 12500
for N in 16 32 64; do python tools/time_stages.py $N $((2000000*16/N)) | grep -E "nofbar|memset" | tee -a gpurun_out/r2f_stage_noload.jsonl; done
echo DONE
