#!/bin/bash
# round 2, call AF (2 GPUs): the NCCL all-gather inside the captured Newton iteration: two-rank test, then the torchrun bench with cfg5
cd "$GRAFT_REPO_ROOT"; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_boundary.py -x -q -m gpu -k "two_ranks or multi_device or cuda_graph" > gpurun_out/r2af_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2af_pytest.log; tail -4 gpurun_out/r2af_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2af_bench_2gpu.json 2> gpurun_out/r2af_bench_2gpu.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2af_bench_2gpu.json; tail -3 gpurun_out/r2af_bench_2gpu.err
SRI_NEWTON_GRAPH=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2af_bench_2gpu_eager.json 2> gpurun_out/r2af_bench_2gpu_eager.err; echo "bench eager rc=$?"
echo DONE
