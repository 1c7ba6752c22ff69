"""Newton static shape solve (native driver, analytic Jacobian, device-side test): one CUDA graph launch per iteration against
eager launches, at the per-GPU batch sizes of the 1- and 8-GPU cfg5 runs."""
import json, os, sys, time
sys.path.insert(0, '.')
import torch
from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator, kernel_launch_count
dev = torch.device("cuda", 0)
for rods in (100000, 25000, 12500, 3000):
    for fd in (0.0, 1e-6):
        row = {"rods": rods, "jacobian": "analytic" if fd == 0.0 else "forward differences"}
        for mode in ("graph", "eager"):
            os.environ["SRI_NEWTON_GRAPH"] = "1" if mode == "graph" else "0"
            with SpectralRodIntegrator(16, 0) as h:
                F = torch.empty((rods, 3), dtype=torch.float64, device=dev)
                h.generate_rods(0x5EED, 0, rods, None, F, None, None)
                F[:, 2] = -(F[:, 2] + 1.0); F[:, :2] = 0.0
                Mt = torch.zeros_like(F)
                ts = []
                for rep_i in range(7):
                    torch.cuda.synchronize()
                    l0 = kernel_launch_count(); t0 = time.perf_counter()
                    qe, rep = h.newton_static_shape(F, Mt, 3, (1.0, 1.0, 0.77), tol=1e-10, max_iter=30, fd_step=fd)
                    torch.cuda.synchronize()
                    ts.append((time.perf_counter() - t0) * 1e3)
                ts = sorted(ts[1:])
                row[mode + "_ms_best"] = ts[0]; row[mode + "_ms_median"] = ts[len(ts) // 2]
                row["iterations"] = rep["iterations"]; row["kernels_per_solve"] = kernel_launch_count() - l0
        print(json.dumps(row))
