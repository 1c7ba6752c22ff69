#!/bin/bash
# round 2, call P: final 1-GPU bench line (both arms) on the final library
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2p_pytest.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2p_bench_reference.json 2> gpurun_out/r2p_bench_reference.err; echo "ref rc=$?"
python bench.py --steps 20 --warmup 5 > gpurun_out/r2p_bench_1gpu.json 2> gpurun_out/r2p_bench_1gpu.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2p_bench_1gpu.json
echo DONE
