"""Throughput of the separate-stage entry points (HBM-bound rows a8-a11 of SURVEY 8) against the HBM roofline."""
import json, sys
sys.path.insert(0, '.')
import torch
from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2000000
M = N - 1
h = SpectralRodIntegrator(N, 0); h.set_stream(torch.cuda.current_stream())
f64 = torch.float64
K = torch.empty((B,3,N), dtype=f64, device='cuda'); F = torch.empty((B,3), dtype=f64, device='cuda'); Mt = torch.empty_like(F); fb = torch.empty_like(K)
h.generate_rods(0x5EED, 0, B, K, F, Mt, fb)
Q = h.integrate_quaternions(K)
r = torch.empty((B,3,M), dtype=f64, device='cuda'); n = torch.empty_like(r); m = torch.empty_like(r)
def timeit(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
cases = {
  "position": (lambda: h.integrate_position(Q, out=r), (4*M + 3*M) * 8),
  "stress_fbar": (lambda: h.integrate_stress(F, fbar=fb, out=n), (3*N + 3 + 3*M) * 8),
  "stress_nofbar": (lambda: h.integrate_stress(F, out=n), (3 + 3*M) * 8),
  "couple": (lambda: h.integrate_couple(Q, n, Mt, out=m), (4*M + 3*M + 3 + 3*M) * 8),
}
for name, (fn, bytes_per_rod) in cases.items():
    ms = timeit(fn)
    print(json.dumps({"stage": name, "N": N, "rods": B, "ms": round(ms, 4), "rods_per_s": B / ms * 1e3, "GBps": B * bytes_per_rod / ms * 1e-6,
                      "frac_of_hbm_6546": B * bytes_per_rod / ms * 1e-6 / 6546.2}))

# calibration of the pure-write roof: a device memset of the same number of bytes as the no-load force stage writes
buf = torch.empty((B * 3 * M,), dtype=f64, device='cuda')
ms = timeit(lambda: buf.zero_())
print(json.dumps({"stage": "memset_same_bytes (cudaMemset-class fill, write-only roof)", "N": N, "rods": B, "ms": round(ms, 4),
                  "GBps": B * 3 * M * 8 / ms * 1e-6, "frac_of_hbm_6546": B * 3 * M * 8 / ms * 1e-6 / 6546.2}))
