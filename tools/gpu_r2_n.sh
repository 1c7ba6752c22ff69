#!/bin/bash
# round 2, call N (4 GPUs): torchrun bench, both arms, as the driver's scaling run does
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 4 --steps 3 --warmup 1 > gpurun_out/r2n_ref_4gpu.json 2> gpurun_out/r2n_ref_4gpu.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/r2n_ref_4gpu.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r2n_bench_4gpu.json 2> gpurun_out/r2n_bench_4gpu.err; echo "bench rc=$?"; tail -c 1200 gpurun_out/r2n_bench_4gpu.json; tail -3 gpurun_out/r2n_bench_4gpu.err
echo DONE
