#!/bin/bash
# round 2, call J (8 GPUs): torchrun bench with the strong-scaling and NCCL Newton legs, C++ multi-GPU example
mkdir -p gpurun_out
nvidia-smi -L | wc -l
examples/multi_gpu_main_gpu 8 1600000
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r2j_bench_8gpu.json 2> gpurun_out/r2j_bench_8gpu.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2j_bench_8gpu.json; tail -5 gpurun_out/r2j_bench_8gpu.err
echo DONE
