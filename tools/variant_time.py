"""Time (and parity-check) one build of libsri_cuda.so selected with SRI_LIB_PATH.  Experiment helper."""
import json, os, sys
sys.path.insert(0, '.')
import numpy as np
import torch
from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
from oracle.oracle import Oracle
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
h = SpectralRodIntegrator(16, 0)
h.set_stream(torch.cuda.current_stream())
o = Oracle(16)
Kc, Fc, Mc, fc = o.generate_rods(0x5EED, 0, 2000)
ref = o.integrate_all(Kc, Fc, Mc, fbar=fc)
t = lambda a: torch.from_numpy(a).cuda()
out = h.integrate_all(t(Kc), t(Fc), t(Mc), fbar=t(fc)); torch.cuda.synchronize()
err = max(float((np.abs(out[s].cpu().numpy() - ref[s]).reshape(2000, -1).max(1) / np.abs(ref[s]).reshape(2000, -1).max(1)).max()) for s in "Qrnm")
dev = 'cuda'
K = torch.empty((B,3,16), dtype=torch.float64, device=dev); F = torch.empty((B,3), dtype=torch.float64, device=dev)
Mt = torch.empty_like(F); fb = torch.empty_like(K)
h.generate_rods(0x5EED, 0, B, K, F, Mt, fb)
Q = torch.empty((B,4,15), dtype=torch.float64, device=dev); r = torch.empty((B,3,15), dtype=torch.float64, device=dev)
n = torch.empty_like(r); m = torch.empty_like(r)
res = {"lib": os.environ.get("SRI_LIB_PATH", "default"), "rods": B, "max_rel_err": err}
for label, kw in (("all4_fbar", dict(fbar=fb, r=r, n=n, m=m)), ("stage1", dict(want=("Q",)))):
    for _ in range(3): h.integrate_all(K, F, Mt, Q=Q, **kw)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): h.integrate_all(K, F, Mt, Q=Q, **kw)
    e1.record(); torch.cuda.synchronize()
    res[label + "_Mrods_s"] = round(B / (e0.elapsed_time(e1) / 10) * 1e-3, 2)
print(json.dumps(res))
