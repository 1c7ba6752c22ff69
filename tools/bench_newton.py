"""BASELINE configs[4]: Newton static shape solve of tip-loaded rods, sharded by rod index, NCCL residual all-reduce.

  python tools/bench_newton.py --rods 100000                      # 1 GPU
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
         tools/bench_newton.py --rods 100000                      # total rods, sharded over the ranks
"""
import argparse, json, os, sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import torch.distributed as dist
from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator, kernel_launch_count
from experimental_gpu_programming_for_a_spectral_numerical_integration_b200.newton import StaticShapeSolver
from experimental_gpu_programming_for_a_spectral_numerical_integration_b200.sharding import shard_range

ap = argparse.ArgumentParser()
ap.add_argument("--rods", type=int, default=100000)
ap.add_argument("--ne", type=int, default=3)
ap.add_argument("--N", type=int, default=16)
ap.add_argument("--jacobian", default="analytic", choices=("analytic", "batched", "columns"), help="analytic: sri_shape_jacobian (one integration per iteration); batched / columns: forward differences")
ap.add_argument("--driver", default="python", choices=("python", "native"), help="newton.py (torch + CUDA graph) or sri_newton_static_shape (loop inside the C ABI)")
args = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
lo, hi = shard_range(args.rods, rank, world)
B = hi - lo
dev = torch.device("cuda", local)
h = SpectralRodIntegrator(args.N, local)
h.set_stream(torch.cuda.current_stream(dev))
# tip force F = (0,0,-f_b), f_b ~ U(0,2) from the rod-index keyed Philox stream (third component of F_tip ~ U(-1,1))
F = torch.empty((B, 3), dtype=torch.float64, device=dev)
h.generate_rods(0x5EED, lo, B, None, F, None, None)
F[:, 2] = -(F[:, 2] + 1.0); F[:, :2] = 0.0
Mt = torch.zeros((B, 3), dtype=torch.float64, device=dev)
from types import SimpleNamespace
from experimental_gpu_programming_for_a_spectral_numerical_integration_b200.sharding import allreduce_residual
if args.driver == "native":
    total_dof = args.rods * 3 * args.ne
    def _allreduce(norms):  # the one collective of the method: 16 bytes per iteration
        t = torch.from_numpy(norms.copy()).to(dev)
        allreduce_residual(t)
        norms[:] = t.cpu().numpy()
    def solve(use_graph=True):
        qe, rep = h.newton_static_shape(F, Mt, args.ne, (1.0, 1.0, 0.77), tol=1e-10, max_iter=30, total_dof=total_dof,
                                        fd_step=0.0 if args.jacobian == "analytic" else 1e-6,
                                        allreduce=_allreduce if world > 1 else None)
        return qe, SimpleNamespace(**rep)
else:
    solver = StaticShapeSolver(h, (1.0, 1.0, 0.77), ne=args.ne, jacobian=args.jacobian)
    def solve(use_graph=True):
        return solver.solve(F, Mt, tol=1e-10, max_iter=30, use_graph=use_graph)
# first solve of this shape (python driver: runs one iteration eagerly, captures the iteration into a CUDA graph, replays it;
# native driver: allocates the workspace); the timed solve below reuses what the first one set up
torch.cuda.synchronize()
t0 = time.perf_counter()
solve()
torch.cuda.synchronize()
first = time.perf_counter() - t0
t0 = time.perf_counter()
solve(use_graph=False)
torch.cuda.synchronize()
eager = time.perf_counter() - t0
if world > 1: dist.barrier()
l0 = kernel_launch_count()
t0 = time.perf_counter()
qe, rep = solve()
torch.cuda.synchronize()
dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
if world > 1: dist.all_reduce(dt, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"config": "cfg5 Newton static shape", "rods_total": args.rods, "n_gpus": world, "N": args.N, "ne": args.ne, "jacobian": args.jacobian, "driver": args.driver,
                      "converged": rep.converged, "newton_iterations": rep.iterations, "integrations_of_the_batch": rep.integrations,
                      "seconds": float(dt.item()), "seconds_eager_launches": eager, "seconds_first_solve_with_capture": first, "rod_solves_per_s": args.rods / float(dt.item()),
                      "rod_integrations_per_s": args.rods * rep.integrations / float(dt.item()),
                      "rms_history": rep.rms_history, "gpu_launches": kernel_launch_count() - l0}))
if world > 1: dist.destroy_process_group()
