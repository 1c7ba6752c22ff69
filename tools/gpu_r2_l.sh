#!/bin/bash
# round 2, call L: full GPU test suite + smoke on the final library
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2l_pytest.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2l_bench.json
echo DONE
