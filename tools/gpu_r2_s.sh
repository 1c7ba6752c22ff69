#!/bin/bash
# round 2, call S: headline kernel with 5 CTAs per SM (96 registers, spills) against the shipped 4 CTAs (128 registers)
mkdir -p gpurun_out
python tools/quick_time.py 1000000 2>&1 | grep all4_fbar | tee gpurun_out/r2s_fused16_occupancy.jsonl
SRI_LIB_PATH=$PWD/tools/_variants/libsri_dmma_mb5.so python tools/quick_time.py 1000000 2>&1 | grep all4_fbar | tee -a gpurun_out/r2s_fused16_occupancy.jsonl
echo DONE
