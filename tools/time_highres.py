import json, sys
sys.path.insert(0, '.')
import torch
from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
for N, B in ((32, 100000), (64, 100000)):
    h = SpectralRodIntegrator(N, 0); h.set_stream(torch.cuda.current_stream())
    M = N - 1
    K = torch.empty((B,3,N), dtype=torch.float64, device='cuda'); F = torch.empty((B,3), dtype=torch.float64, device='cuda')
    Mt = torch.empty_like(F); fb = torch.empty_like(K)
    h.generate_rods(0x5EED, 0, B, K, F, Mt, fb)
    Q = torch.empty((B,4,M), dtype=torch.float64, device='cuda'); r = torch.empty((B,3,M), dtype=torch.float64, device='cuda')
    n = torch.empty_like(r); m = torch.empty_like(r)
    for _ in range(2): h.integrate_all(K, F, Mt, fbar=fb, Q=Q, r=r, n=n, m=m)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): h.integrate_all(K, F, Mt, fbar=fb, Q=Q, r=r, n=n, m=m)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    flops = {32: 1320063, 64: 10869012}[N]
    print(json.dumps({"N": N, "rods": B, "ms": ms, "rods_per_s": B / ms * 1e3, "dense_count_tflops": B / ms * 1e3 * flops * 1e-12}))
    h.close()
