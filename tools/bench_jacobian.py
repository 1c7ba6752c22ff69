"""Times sri_shape_jacobian alone (CUDA events on the handle's stream): 10^5 rods, N = 16, ne = 3 by default.
  python tools/bench_jacobian.py [rods] [ne]        (SRI_LIB_PATH / SRI_JACOBIAN_IMPL=scalar select the build / the scalar kernel)"""
import json
import os
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
ne = int(sys.argv[2]) if len(sys.argv) > 2 else 3
h = SpectralRodIntegrator(16, 0)
st = torch.cuda.current_stream()
h.set_stream(st)
f64 = torch.float64
K = torch.empty((B, 3, 16), dtype=f64, device="cuda"); F = torch.empty((B, 3), dtype=f64, device="cuda"); M = torch.empty((B, 3), dtype=f64, device="cuda")
h.generate_rods(0x5EED, 0, B, K, F, M, None)
out = h.integrate_all(K, F, M)
J = torch.empty((B, 3 * ne, 3 * ne), dtype=f64, device="cuda")
run = lambda: h.shape_jacobian(out["Q"], out["n"], out["m"], M, ne, (1.0, 1.0, 0.77), out=J)
for _ in range(3):
    run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 20
e0.record(st)
for _ in range(reps):
    run()
e1.record(st)
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / reps * 1e3
nt = (9 * ne + 7) // 8
print(json.dumps({"kernel": "sri_shape_jacobian", "impl": os.environ.get("SRI_JACOBIAN_IMPL", "dmma"), "lib": os.environ.get("SRI_LIB_PATH", "in-tree"),
                  "rods": B, "ne": ne, "us_per_launch": us, "rods_per_s": B / us * 1e6, "dmma_per_rod": 20 * nt,
                  "dmma_tflops": 20 * nt * 512 * B / us * 1e-6, "checksum": float(J.sum().item())}))
