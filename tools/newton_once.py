"""One analytic-Jacobian Newton solve of 10^5 tip-loaded rods (bench.py's cfg5 workload) -- the command the ncu launch list of
the Newton driver is taken on:  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file L python tools/newton_once.py"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch  # noqa: E402

from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
fd = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
h = SpectralRodIntegrator(16, 0)
F = torch.empty((B, 3), dtype=torch.float64, device="cuda")
h.generate_rods(0x5EED, 0, B, None, F, None, None)
F[:, 2] = -(F[:, 2] + 1.0)
F[:, :2] = 0.0
M = torch.zeros_like(F)
h.newton_static_shape(F, M, 3, (1.0, 1.0, 0.77), fd_step=fd)
torch.cuda.synchronize()
t0 = time.perf_counter()
_, rep = h.newton_static_shape(F, M, 3, (1.0, 1.0, 0.77), fd_step=fd)
torch.cuda.synchronize()
print(f"{B} rods: {1e3 * (time.perf_counter() - t0):.3f} ms, {rep['iterations']} iterations, rms {rep['rms']:.2e}")
