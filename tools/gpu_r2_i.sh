#!/bin/bash
# round 2, call I (2 GPUs): multi-device tests, C++ multi-GPU example, torchrun bench with the strong-scaling and NCCL Newton legs
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_boundary.py -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2i_pytest.log
examples/multi_gpu_main_gpu 2 400000
NCCL_DEBUG=WARN python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2i_bench_2gpu.json 2> gpurun_out/r2i_bench_2gpu.err; echo "bench rc=$?"; tail -c 5000 gpurun_out/r2i_bench_2gpu.json; tail -5 gpurun_out/r2i_bench_2gpu.err
echo DONE
