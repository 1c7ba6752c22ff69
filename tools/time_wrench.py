"""Throughput of the direct local-frame statics solve (sri_integrate_wrench_local) and of its pointwise form."""
import json, sys
sys.path.insert(0, '.')
import torch
from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 16
h = SpectralRodIntegrator(N, 0); h.set_stream(torch.cuda.current_stream())
f64 = torch.float64
K = torch.empty((B,3,N), dtype=f64, device='cuda'); F = torch.empty((B,3), dtype=f64, device='cuda'); Mt = torch.empty_like(F); fb = torch.empty_like(K)
h.generate_rods(0x5EED, 0, B, K, F, Mt, fb)
out = h.integrate_all(K, F, Mt, fbar=fb)
lam = torch.empty((B,6,N), dtype=f64, device='cuda')
def timeit(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
ms = timeit(lambda: h.integrate_wrench_local(K, out["Q"], F, Mt, fbar=fb, out=lam))
hb = h.handback_count()
n3 = 3 * (N - 1)
print(json.dumps({"op": f"sri_integrate_wrench_local ({n3}x{n3} LU + 2 solves per rod)", "N": N, "rods": B, "ms": ms, "rods_per_s": B / ms * 1e3, "handed_back": hb,
                  "gflops_lu": B / ms * 1e3 * (2 * n3 ** 3 / 3 + 4 * n3 ** 2) * 1e-9}))
ms = timeit(lambda: h.wrench_local(out["Q"], out["n"], out["m"], F, Mt, out=lam))
print(json.dumps({"op": "sri_wrench_local (pointwise)", "N": N, "rods": B, "ms": ms, "rods_per_s": B / ms * 1e3}))
