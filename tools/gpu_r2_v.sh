#!/bin/bash
# round 2, run V: multi-warp Gauss-Jordan for the local-frame statics at 17 <= N <= 33: parity, then throughput against the CTA-wide LU
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "local_frame" > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
for N in 17 20 22 24 32 33; do
  timeout 120 python tools/time_wrench.py 20000 $N | head -1 >> gpurun_out/r2v_wrench.jsonl 2>&1
  SRI_WRENCH_IMPL=generic timeout 120 python tools/time_wrench.py 20000 $N | head -1 | sed 's/^/generic /' >> gpurun_out/r2v_wrench.jsonl 2>&1
done
timeout 120 python tools/time_wrench.py 200000 16 | head -1 >> gpurun_out/r2v_wrench.jsonl 2>&1
tail -5 gpurun_out/r2v_pytest.log; cat gpurun_out/r2v_wrench.jsonl
