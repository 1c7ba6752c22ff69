#!/bin/bash
# round 2, call M (2 GPUs): boundary tests incl. the two-rank NCCL test
mkdir -p gpurun_out
python -m pytest tests/test_gpu_boundary.py -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2m_pytest.log
echo DONE
