#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_newton.py -m gpu -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2t_pytest.log
echo DONE
