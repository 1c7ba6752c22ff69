"""Critical-path probe: time the fused kernel when each SM sub-partition holds exactly 1, 2, 3 warps (one iteration)."""
import json, os, sys
sys.path.insert(0, '.')
import torch
from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
h = SpectralRodIntegrator(16, 0)
h.set_stream(torch.cuda.current_stream())
dev = 'cuda'
res = {"lib": os.environ.get("SRI_LIB_PATH", "default")}
for wpb in (1, 2, 3, 6, 12, 24):
    B = 148 * 4 * 2 * wpb   # wpb warps per SM sub-partition in total (resident up to occupancy, then serial)
    K = torch.empty((B,3,16), dtype=torch.float64, device=dev); F = torch.empty((B,3), dtype=torch.float64, device=dev)
    Mt = torch.empty_like(F); fb = torch.empty_like(K)
    h.generate_rods(0x5EED, 0, B, K, F, Mt, fb)
    Q = torch.empty((B,4,15), dtype=torch.float64, device=dev); r = torch.empty((B,3,15), dtype=torch.float64, device=dev)
    n = torch.empty_like(r); m = torch.empty_like(r)
    for label, kw in (("all4", dict(fbar=fb, r=r, n=n, m=m)), ("stage1", dict(want=("Q",)))):
        for _ in range(5): h.integrate_all(K, F, Mt, Q=Q, **kw)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(20):
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); h.integrate_all(K, F, Mt, Q=Q, **kw); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        res[f"{label}_w{wpb}_us"] = round(best * 1e3, 2)
print(json.dumps(res))
