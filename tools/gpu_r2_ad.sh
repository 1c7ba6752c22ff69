#!/bin/bash
# round 2, run AD: local-frame statics after restricting the static-order first pass to N = 17, 23..33: parity (all wrench tests), throughput per N
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "local_frame" > gpurun_out/r2ad_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ad_pytest.log
for N in 5 8 11 12 16; do timeout 120 python tools/time_wrench.py 200000 $N 2>/dev/null | head -1 >> gpurun_out/r2ad_wrench.jsonl; done
for N in 17 18 20 22 23 32 33; do timeout 120 python tools/time_wrench.py 40000 $N 2>/dev/null | head -1 >> gpurun_out/r2ad_wrench.jsonl; done
for N in 34 40 64; do timeout 120 python tools/time_wrench.py 4000 $N 2>/dev/null | head -1 >> gpurun_out/r2ad_wrench.jsonl; done
grep "rc=\|passed\|failed\|Error\|assert" gpurun_out/r2ad_pytest.log | head -20; cut -c1-230 gpurun_out/r2ad_wrench.jsonl
