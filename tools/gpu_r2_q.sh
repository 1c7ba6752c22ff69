#!/bin/bash
# round 2, call Q: register-resident wrench kernel -- parity tests and A/B timing against the blocked shared-memory kernel
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "local_frame or wrench" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2q_pytest.log
python tools/time_wrench.py 200000 16 | tee gpurun_out/r2q_wrench.jsonl
SRI_WRENCH_IMPL=blocked python tools/time_wrench.py 200000 16 | tee -a gpurun_out/r2q_wrench.jsonl
python tools/time_wrench.py 200000 12 | tee -a gpurun_out/r2q_wrench.jsonl
echo DONE
