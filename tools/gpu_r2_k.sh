#!/bin/bash
# round 2, call K: evidence for profiles/ -- tests, both bench arms, launch lists, ncu captures of the two hot kernels
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2k_pytest.log
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2k_bench_reference.json 2> gpurun_out/r2k_bench_reference.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r2k_bench_reference.json
python bench.py --steps 20 --warmup 5 > gpurun_out/r2k_bench_1gpu.json 2> gpurun_out/r2k_bench_1gpu.err; echo "bench rc=$?"; tail -c 2500 gpurun_out/r2k_bench_1gpu.json; tail -3 gpurun_out/r2k_bench_1gpu.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-side-configs --rods 200000"
$CMD > gpurun_out/r2k_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2k_launches.csv $CMD > gpurun_out/r2k_ncu_launches.log 2>&1
$CMD > gpurun_out/r2k_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fused16_dmma -s 3 -c 1 -o gpurun_out/r2k_prof_fused16_dmma $CMD > gpurun_out/r2k_ncu_full.log 2>&1; tail -2 gpurun_out/r2k_ncu_full.log
python tools/bench_jacobian.py | tee gpurun_out/r2k_jacobian.jsonl
ncu --set full --clock-control none --import-source on -k regex:shape_jacobian_dmma -s 2 -c 1 -o gpurun_out/r2k_prof_jacobian python tools/bench_jacobian.py > gpurun_out/r2k_ncu_jac.log 2>&1; tail -2 gpurun_out/r2k_ncu_jac.log
python tools/newton_once.py && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2k_newton_launches.csv python tools/newton_once.py > gpurun_out/r2k_ncu_newton.log 2>&1
python tools/time_stages.py 16 2000000 | tee gpurun_out/r2k_stage_kernels.jsonl
python tools/time_stages.py 32 400000 | tee -a gpurun_out/r2k_stage_kernels.jsonl
python tools/time_stages.py 64 200000 | tee -a gpurun_out/r2k_stage_kernels.jsonl
python tools/time_wrench.py 200000 16 | tee gpurun_out/r2k_wrench.jsonl
python tools/time_wrench.py 20000 32 | tee -a gpurun_out/r2k_wrench.jsonl
python tools/time_wrench.py 4000 64 | tee -a gpurun_out/r2k_wrench.jsonl
echo DONE
