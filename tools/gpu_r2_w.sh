#!/bin/bash
# round 2, run W: one-row-per-lane Gauss-Jordan (SRI_WRENCH_IMPL=multi) against the two-rows-per-lane single-warp kernel at N <= 17
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
SRI_WRENCH_IMPL=multi timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "local_frame" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2w_pytest.log
for N in 16 12 11 8 17; do
  timeout 120 python tools/time_wrench.py 200000 $N 2>/dev/null | head -1 >> gpurun_out/r2w_wrench.jsonl
  SRI_WRENCH_IMPL=multi timeout 120 python tools/time_wrench.py 200000 $N 2>/dev/null | head -1 | sed 's/^/multi /' >> gpurun_out/r2w_wrench.jsonl
done
tail -5 gpurun_out/r2w_pytest.log; cut -c1-200 gpurun_out/r2w_wrench.jsonl
