#!/bin/bash
# round 2, call B: GPU tests (all), Newton launch list
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/r2b_pytest.log
python tools/newton_once.py && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_newton_launches.csv python tools/newton_once.py > gpurun_out/r2b_ncu_newton.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2b_newton_launches.csv')) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = r[4][:60]; v = float(r[-1].replace(',', ''))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in agg.items(): print(f"{n:4d} x {k:62s} {t/1e3:10.1f} us total")
PY
echo DONE
