#!/bin/bash
# round 2, call C: GPU tests, Jacobian A/B, Newton timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/r2c_pytest.log
for v in "" "tools/_variants/libsri_jac4.so"; do
  echo "== variant '$v'"
  SRI_LIB_PATH=${v:+$PWD/$v} python tools/bench_jacobian.py | tee -a gpurun_out/r2c_jacobian.jsonl
  SRI_LIB_PATH=${v:+$PWD/$v} python tools/bench_jacobian.py 12500 | tee -a gpurun_out/r2c_jacobian.jsonl
  SRI_LIB_PATH=${v:+$PWD/$v} python tools/newton_once.py
  SRI_LIB_PATH=${v:+$PWD/$v} python tools/newton_once.py 12500
done
SRI_JACOBIAN_IMPL=scalar python tools/bench_jacobian.py | tee -a gpurun_out/r2c_jacobian.jsonl
python tools/bench_jacobian.py 100000 5 | tee -a gpurun_out/r2c_jacobian.jsonl
echo DONE
