#!/bin/bash
# round 2, call A: GPU tests + 1-GPU bench (both arms)
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2a_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/r2a_bench.json; tail -5 gpurun_out/r2a_bench.err
