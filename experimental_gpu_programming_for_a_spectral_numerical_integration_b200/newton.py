"""Static shape solve of tip-loaded rods (BASELINE configs[4]; rod_modeling.pdf section 2.2).

Unknowns: modal strain coordinates qe [batch][3*ne] (Legendre modes, the reference's Phi of include/utilities.h:49-67).
Residual: Galerkin projection of the constitutive mismatch (rod_modeling.pdf eq. 1.25 projected as in 2.14/2.20)
    g(qe) = int_0^1 Phi^T ( H (K - K0) - R(q)^T m ) dX,        K = Phi qe,
where q, m come from the four-stage integration of K under the tip wrench (F_tip, M_tip).  Newton iteration per rod
with a forward-difference Jacobian: every column costs one fused four-stage integration of the whole batch, so one
iteration is 3*ne+1 integrations of the batch -- this driver is the hot path's main caller.  The 3*ne perturbed copies of
the batch are integrated by ONE call of the hot path on 3*ne*B rods (jacobian="batched", the default: four library
launches and four elementwise torch kernels per Jacobian, full waves on every SM); jacobian="columns" integrates them
one after the other on the B-rod buffers (3*ne times the launches, 1/(3*ne) of the workspace).  jacobian="analytic" takes
the Jacobian from sri_shape_jacobian instead (left-trivialised rotation variation: two contractions per direction, no
integration): one integration of the batch per iteration.

All arithmetic runs in this repository's CUDA kernels through the C ABI; torch supplies buffers, the trivial
axpy-style updates of qe and the process group.  Multi-GPU: rods are sharded by index, the only collective is the
all-reduce of [sum g^2, max |g|] per iteration (sharding.allreduce_residual).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import torch

from .api import SpectralRodIntegrator
from .sharding import allreduce_residual


@dataclass
class NewtonReport:
    iterations: int
    converged: bool
    rms_history: List[float] = field(default_factory=list)
    max_history: List[float] = field(default_factory=list)
    integrations: int = 0


class StaticShapeSolver:
    def __init__(self, integrator: SpectralRodIntegrator, H_diag=(1.0, 1.0, 0.77), ne: int = 4, fd_step: float = 1e-6,
                 jacobian: str = "batched"):
        if not 1 <= ne <= 8:
            raise ValueError("1 <= ne <= 8")
        if jacobian not in ("batched", "columns", "analytic"):
            raise ValueError('jacobian is "batched", "columns" or "analytic"')
        self.jacobian = jacobian
        self.h = integrator
        self.H = tuple(float(v) for v in H_diag)
        self.ne = int(ne)
        self.fd_step = float(fd_step)
        self._cache = {}

    # -- one evaluation of g(qe): strain samples, fused four-stage integration (+ its second pass), Galerkin residual
    def residual(self, qe, F_tip, M_tip, K0=None, work=None, out=None, reduce=None):
        h = self.h
        K = h.strain_from_modes(qe, out=None if work is None else work["K"])
        keep_n = work is not None and work.get("n") is not None   # the analytic Jacobian needs the internal force too
        res = h.integrate_all(K, F_tip, M_tip, Q=None if work is None else work["Q"], m=None if work is None else work["m"],
                              n=work["n"] if keep_n else None, want=("Q", "n", "m") if keep_n else ("Q", "m"))
        if out is None and work is not None:
            out = work["g"]
        return h.galerkin_residual(K, self.H, res["Q"], res["m"], M_tip, self.ne, K0=K0, out=out, reduce=reduce)

    # -- persistent buffers (and the captured iteration) for one problem shape
    def _workspace(self, B: int, dev, has_K0: bool):
        key = (B, dev.index, has_K0)
        ws = self._cache.get(key)
        if ws is None:
            h, n, f64 = self.h, 3 * self.ne, torch.float64
            N, M = h.N, h.M
            e = lambda *shape: torch.empty(shape, dtype=f64, device=dev)
            self._cache.clear()  # one shape at a time: the buffers of a 10^6-rod problem are not small
            ws = {"K": e(B, 3, N), "Q": e(B, 4, M), "m": e(B, 3, M), "g": e(B, n),
                  "F": e(B, 3), "Mt": e(B, 3), "K0": e(B, 3, N) if has_K0 else None,
                  "qe": e(B, n), "g0": e(B, n), "J": e(B, n, n), "qp": e(B, n), "delta": e(B, n),
                  "red": torch.zeros(2, dtype=f64, device=dev), "graph": None, "warmed": False, "wide": None,
                  "n": e(B, 3, M) if self.jacobian == "analytic" else None}
            if self.jacobian == "batched":  # the n perturbed copies of the batch, copy d = rods [d B, (d+1) B)
                ws["wide"] = {"qe": e(n, B, n), "K": e(n * B, 3, N), "Q": e(n * B, 4, M), "m": e(n * B, 3, M),
                              "g": e(n * B, n), "F": e(n, B, 3), "Mt": e(n, B, 3),
                              "K0": e(n, B, 3, N) if has_K0 else None}
            self._cache[key] = ws
        return ws

    def _evaluate(self, ws):
        """g0 <- g(qe) and red <- [sum g0^2, max |g0|] (this rank's rods)."""
        if ws["qe"].shape[0] == 0:
            ws["red"].zero_()
            return
        self.residual(ws["qe"], ws["F"], ws["Mt"], ws["K0"], ws, out=ws["g0"], reduce=ws["red"])

    def _iteration(self, ws):
        """One Newton iteration on static buffers only (capturable): forward-difference Jacobian (3 ne integrations of
        the whole batch), batched per-rod solve, update, residual of the new iterate."""
        qe, qp, g0, J = ws["qe"], ws["qp"], ws["g0"], ws["J"]
        n, B, wide = 3 * self.ne, ws["qe"].shape[0], ws["wide"]
        if self.jacobian == "analytic":   # Q, n, m of the current iterate are in the workspace (last _evaluate)
            self.h.shape_jacobian(ws["Q"], ws["n"], ws["m"], ws["Mt"], self.ne, self.H, out=J)
        elif wide is not None:
            wq = wide["qe"]
            wq.copy_(qe.unsqueeze(0))
            wq.diagonal(dim1=0, dim2=2).add_(self.fd_step)            # copy d: qe + fd_step e_d
            gw = self.residual(wq.view(n * B, n), wide["F"].view(n * B, 3), wide["Mt"].view(n * B, 3),
                               None if wide["K0"] is None else wide["K0"].view(n * B, 3, self.h.N), wide)
            J.copy_(((gw.view(n, B, n) - g0.unsqueeze(0)) / self.fd_step).permute(1, 2, 0))
        else:
            for d in range(n):
                qp.copy_(qe)
                qp[:, d] += self.fd_step
                gd = self.residual(qp, ws["F"], ws["Mt"], ws["K0"], ws)
                J[:, :, d] = (gd - g0) / self.fd_step
        self.h.solve_small_batched(J, g0, out=ws["delta"])  # per-rod n x n Newton system
        qe -= ws["delta"]
        self._evaluate(ws)

    def _capture(self, ws, dev):
        """Capture one iteration into a CUDA graph (jacobian="columns": ~12 launches per integration, 3 ne + 1
        integrations; launch-bound at 10^5 rods per GPU, 0.44 ms of kernel per integration against ~1.2 ms of launches)."""
        h = self.h
        was_explicit, was_stream = h._explicit_stream, torch.cuda.current_stream(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            h.set_stream(torch.cuda.current_stream(dev))  # this library's launches join the capture
            self._iteration(ws)
        h.set_stream(was_stream if was_explicit else None)
        ws["graph"] = graph

    def solve(self, F_tip, M_tip, qe0=None, K0=None, tol: float = 1e-10, max_iter: int = 30, group=None,
              use_graph: bool = True):
        """Newton iteration until the GLOBAL rms of g over all ranks' rods is below tol.  Returns (qe, NewtonReport).

        The first iteration of a new problem shape runs eagerly, the second is captured into a CUDA graph, every later
        one -- and every iteration of later solve() calls with the same batch size -- replays it."""
        ne = self.ne
        B = F_tip.shape[0]
        n = 3 * ne
        dev, f64 = F_tip.device, torch.float64
        ws = self._workspace(B, dev, K0 is not None)
        ws["F"].copy_(F_tip)
        ws["Mt"].copy_(M_tip)
        if K0 is not None:
            ws["K0"].copy_(K0)
        if ws["wide"] is not None:  # the tip loads (and K0) of the perturbed copies do not change over the iterations
            ws["wide"]["F"].copy_(ws["F"].unsqueeze(0))
            ws["wide"]["Mt"].copy_(ws["Mt"].unsqueeze(0))
            if K0 is not None:
                ws["wide"]["K0"].copy_(ws["K0"].unsqueeze(0))
        if qe0 is None:
            ws["qe"].zero_()
        else:
            ws["qe"].copy_(qe0)
        red = ws["red"]
        count = torch.tensor([float(B * n)], dtype=f64, device=dev)
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            torch.distributed.all_reduce(count, group=group)
        total_dof = float(count.item())
        rep = NewtonReport(iterations=0, converged=False)
        self._evaluate(ws)
        rep.integrations += 1
        for it in range(max_iter + 1):
            allreduce_residual(red, group)                     # the only collective: 16 bytes per iteration
            sum2, gmax = red.tolist()                          # one device-to-host read per iteration
            rms = math.sqrt(sum2 / total_dof) if sum2 == sum2 else float("nan")
            rep.rms_history.append(rms)
            rep.max_history.append(gmax)
            if rms < tol:
                rep.converged = True
                break
            if it == max_iter:
                break
            if use_graph and B > 0 and ws["graph"] is None and ws["warmed"]:
                self._capture(ws, dev)
            if use_graph and ws["graph"] is not None:
                ws["graph"].replay()
            else:
                self._iteration(ws)
                ws["warmed"] = True
            rep.iterations += 1
            rep.integrations += 1 if self.jacobian == "analytic" else n + 1
        return ws["qe"].clone(), rep
