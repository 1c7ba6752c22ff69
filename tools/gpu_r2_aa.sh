#!/bin/bash
# round 2, run AA: static-order first pass of the local-frame statics (preconditioned Gauss-Jordan + hand-back): parity, then throughput per N
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "local_frame" > gpurun_out/r2aa_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2aa_pytest.log
SRI_DMMA_GROWTH=0 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "local_frame" >> gpurun_out/r2aa_pytest.log 2>&1; echo "pytest growth0 rc=$?" >> gpurun_out/r2aa_pytest.log
for N in 5 8 11 12 16; do timeout 120 python tools/time_wrench.py 200000 $N 2>/dev/null | head -1 >> gpurun_out/r2aa_wrench.jsonl; done
SRI_WRENCH_STATIC16=0 timeout 120 python tools/time_wrench.py 200000 16 2>/dev/null | head -1 | sed 's/^/static16=0 /' >> gpurun_out/r2aa_wrench.jsonl
for N in 17 18 22 23 32 33; do timeout 120 python tools/time_wrench.py 40000 $N 2>/dev/null | head -1 >> gpurun_out/r2aa_wrench.jsonl; done
grep "rc=\|passed\|failed\|Error\|assert" gpurun_out/r2aa_pytest.log | head -20; cut -c1-200 gpurun_out/r2aa_wrench.jsonl
