#!/bin/bash
# full ncu capture of the N=32 and N=64 DMMA kernels
mkdir -p gpurun_out
cat > /tmp/hr.py <<'PY'
import sys
sys.path.insert(0, '.')
import torch
from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
N, B = int(sys.argv[1]), int(sys.argv[2])
h = SpectralRodIntegrator(N, 0); h.set_stream(torch.cuda.current_stream()); M = N - 1
K = torch.empty((B,3,N), dtype=torch.float64, device='cuda'); F = torch.empty((B,3), dtype=torch.float64, device='cuda')
Mt = torch.empty_like(F); fb = torch.empty_like(K)
h.generate_rods(0x5EED, 0, B, K, F, Mt, fb)
for _ in range(3): h.integrate_all(K, F, Mt, fbar=fb)
torch.cuda.synchronize()
PY
python /tmp/hr.py 32 40000 && ncu --set full --clock-control none --import-source on -k regex:tiled_dmma -s 1 -c 1 -o gpurun_out/prof_dmma32 python /tmp/hr.py 32 40000 > gpurun_out/ncu32.log 2>&1
python /tmp/hr.py 64 8000 && ncu --set full --clock-control none --import-source on -k regex:tiled_dmma -s 1 -c 1 -o gpurun_out/prof_dmma64 python /tmp/hr.py 64 8000 > gpurun_out/ncu64.log 2>&1
tail -2 gpurun_out/ncu32.log gpurun_out/ncu64.log
