#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r2u_bench_1gpu.json 2> gpurun_out/r2u_bench_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/r2u_bench_1gpu.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2u_bench_1gpu.json').read().strip().split("\n")[-1])
print(json.dumps(d["other_configs"]["stages"], indent=1)[:2500])
print(d["value"], d["other_configs"]["cfg5"]["analytic_jacobian"]["seconds"])
PY
echo DONE
