import sys, time, json
sys.path.insert(0, '.')
import torch
from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
h = SpectralRodIntegrator(16, 0)
h.use_current_torch_stream()
dev = 'cuda'
K = torch.empty((B,3,16), dtype=torch.float64, device=dev); F = torch.empty((B,3), dtype=torch.float64, device=dev)
Mt = torch.empty_like(F); fb = torch.empty_like(K)
h.generate_rods(0x5EED, 0, B, K, F, Mt, fb)
Q = torch.empty((B,4,15), dtype=torch.float64, device=dev); r = torch.empty((B,3,15), dtype=torch.float64, device=dev)
n = torch.empty_like(r); m = torch.empty_like(r)
print("fp64 peak TF:", h.measure_fp64_peak())
for label, kw in (("all4_fbar", dict(fbar=fb)), ("all4_nofbar", dict()),):
    for _ in range(3):
        h.integrate_all(K, F, Mt, Q=Q, r=r, n=n, m=m, **kw)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        h.integrate_all(K, F, Mt, Q=Q, r=r, n=n, m=m, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"case": label, "rods": B, "ms": ms, "rods_per_s": B / ms * 1e3}))
# stage-1 only
for _ in range(3): h.integrate_quaternions(K, out=Q)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): h.integrate_quaternions(K, out=Q)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
print(json.dumps({"case": "stage1_only", "rods": B, "ms": ms, "rods_per_s": B / ms * 1e3}))
