#!/bin/bash
# round 2, call G: tiled DMMA kernel with the reciprocal chain in the shadow of the normalisation DMMAs (A/B), final Jacobian
mkdir -p gpurun_out
python tools/time_highres.py | tee -a gpurun_out/r2g_highres.jsonl
SRI_LIB_PATH=$PWD/tools/_variants/libsri_tiled_reorder.so python tools/time_highres.py | tee -a gpurun_out/r2g_highres.jsonl
python tools/time_highres.py | tee -a gpurun_out/r2g_highres.jsonl
SRI_LIB_PATH=$PWD/tools/_variants/libsri_tiled_reorder.so python tools/time_highres.py | tee -a gpurun_out/r2g_highres.jsonl
python tools/bench_jacobian.py | tee -a gpurun_out/r2g_jacobian.jsonl
python tools/newton_once.py; python tools/newton_once.py 12500
echo DONE
