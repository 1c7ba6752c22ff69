#!/bin/bash
# round 2, call H: diagnostic timing of the tiled DMMA kernel without the owner's chain / without prepare (results are wrong on purpose)
mkdir -p gpurun_out
export SRI_DMMA_GROWTH=1e300
for v in NOCHAIN NOPREPARE; do echo "== $v"; SRI_LIB_PATH=$PWD/tools/_variants/libsri_diag_$v.so python tools/time_highres.py | tee -a gpurun_out/r2h_tiled_diag_$v.jsonl; done
unset SRI_DMMA_GROWTH
python tools/time_highres.py | tee -a gpurun_out/r2h_tiled_ref.jsonl
echo DONE
