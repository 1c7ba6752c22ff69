#!/bin/bash
# round 2, call D: GPU tests, Jacobian A/B + ncu, small solve, stage kernels, wrench N = 16/32/64
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2d_pytest.log
python tools/bench_jacobian.py | tee -a gpurun_out/r2d_jacobian.jsonl
SRI_LIB_PATH=$PWD/tools/_variants/libsri_jac3.so python tools/bench_jacobian.py | tee -a gpurun_out/r2d_jacobian.jsonl
python tools/newton_once.py; python tools/newton_once.py 12500
python tools/newton_once.py && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2d_newton_launches.csv python tools/newton_once.py > gpurun_out/r2d_ncu_newton.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2d_newton_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4][:60]; v = float(r[-1].replace(',', ''))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items(): print(f"{n:4d} x {k:62s} {t/1e3:10.1f} us total")
PY
for N in 16 32 64; do python tools/time_stages.py $N $((2000000*16/N)) | grep nofbar | tee -a gpurun_out/r2d_stage_noload.jsonl; done
python tools/time_wrench.py 200000 16 | tee -a gpurun_out/r2d_wrench.jsonl
python tools/time_wrench.py 20000 32 | tee -a gpurun_out/r2d_wrench.jsonl
python tools/time_wrench.py 4000 64 | tee -a gpurun_out/r2d_wrench.jsonl
ncu --set full --clock-control none --import-source on -k regex:shape_jacobian_dmma -s 2 -c 1 -o gpurun_out/r2d_prof_jacobian python tools/bench_jacobian.py > gpurun_out/r2d_ncu_jac.log 2>&1; tail -2 gpurun_out/r2d_ncu_jac.log
echo DONE
