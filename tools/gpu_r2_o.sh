#!/bin/bash
# round 2, call O: persistent Galerkin residual kernel -- Newton tests + timing + launch list
mkdir -p gpurun_out
python -m pytest tests/test_gpu_newton.py tests/test_gpu_boundary.py -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2o_pytest.log
python tools/newton_once.py; python tools/newton_once.py; python tools/newton_once.py 12500
python tools/newton_once.py 100000 1e-6
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2o_newton_launches.csv python tools/newton_once.py > gpurun_out/r2o_ncu_newton.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2o_newton_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4][:60]; v = float(r[-1].replace(',', ''))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items(): print(f"{n:4d} x {k:62s} {t/1e3:10.1f} us total")
PY
echo DONE
