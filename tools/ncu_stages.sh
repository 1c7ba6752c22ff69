#!/bin/bash
# ncu counters of the separate-stage kernels (a handful of metrics, a few launches each)
mkdir -p gpurun_out
MET=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,launch__registers_per_thread
ncu --metrics $MET --clock-control none -k regex:"stage_tma|stage_dmma" -s 9 -c 9 --csv --log-file gpurun_out/stages16_ncu.csv python tools/time_stages.py 16 400000 > /dev/null 2>&1
ncu --metrics $MET --clock-control none -k regex:stage_generic -s 9 -c 9 --csv --log-file gpurun_out/stages64_ncu.csv python tools/time_stages.py 64 100000 > /dev/null 2>&1
ncu --metrics $MET --clock-control none -k regex:stage_generic -s 9 -c 9 --csv --log-file gpurun_out/stages32_ncu.csv python tools/time_stages.py 32 200000 > /dev/null 2>&1
wc -l gpurun_out/stages*_ncu.csv
