#!/bin/bash
# round 2, run AE: Newton iteration as a CUDA graph: tests, then graph against eager launches at several batch sizes
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_boundary.py tests/test_gpu_newton.py -x -q -m gpu > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2ae_pytest.log
timeout 600 python tools/time_newton_graph.py > gpurun_out/r2ae_newton_graph.jsonl 2>gpurun_out/r2ae_err.log
tail -5 gpurun_out/r2ae_pytest.log; cat gpurun_out/r2ae_newton_graph.jsonl; tail -3 gpurun_out/r2ae_err.log
