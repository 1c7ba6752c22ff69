#!/bin/bash
# round 2, run Y: local-frame statics, default dispatch after the one-row-per-lane kernels: parity (all wrench tests, default and forced variants), throughput per N
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "local_frame" > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
SRI_WRENCH_IMPL=multi timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "local_frame" >> gpurun_out/r2y_pytest.log 2>&1; echo "pytest multi rc=$?" >> gpurun_out/r2y_pytest.log
SRI_WRENCH_IMPL=warp timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "local_frame" >> gpurun_out/r2y_pytest.log 2>&1; echo "pytest warp rc=$?" >> gpurun_out/r2y_pytest.log
for N in 5 8 11 12 16; do timeout 120 python tools/time_wrench.py 200000 $N 2>/dev/null | head -1 >> gpurun_out/r2y_wrench.jsonl; done
for N in 17 18 20 22 23 32 33; do timeout 120 python tools/time_wrench.py 40000 $N 2>/dev/null | head -1 >> gpurun_out/r2y_wrench.jsonl; done
for N in 40 64; do timeout 120 python tools/time_wrench.py 4000 $N 2>/dev/null | head -1 >> gpurun_out/r2y_wrench.jsonl; done
grep "rc=\|passed\|failed" gpurun_out/r2y_pytest.log; cut -c1-200 gpurun_out/r2y_wrench.jsonl
