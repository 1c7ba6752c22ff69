"""Static shape solve of tip-loaded rods (BASELINE configs[4]; rod_modeling.pdf section 2.2).

Unknowns: modal strain coordinates qe [batch][3*ne] (Legendre modes, the reference's Phi of include/utilities.h:49-67).
Residual: Galerkin projection of the constitutive mismatch (rod_modeling.pdf eq. 1.25 projected as in 2.14/2.20)
    g(qe) = int_0^1 Phi^T ( H (K - K0) - R(q)^T m ) dX,        K = Phi qe,
where q, m come from the four-stage integration of K under the tip wrench (F_tip, M_tip).  Newton iteration per rod
with a forward-difference Jacobian: every column costs one fused four-stage integration of the whole batch, so one
iteration is 3*ne+1 launches of the hot path -- this driver is the hot path's main caller.

All arithmetic runs in this repository's CUDA kernels through the C ABI; torch supplies buffers, the trivial
axpy-style updates of qe and the process group.  Multi-GPU: rods are sharded by index, the only collective is the
all-reduce of [sum g^2, max |g|] per iteration (sharding.allreduce_residual).
"""
from __future__ import annotations

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
    def __init__(self, integrator: SpectralRodIntegrator, H_diag=(1.0, 1.0, 0.77), ne: int = 4, fd_step: float = 1e-6):
        if not 1 <= ne <= 8:
            raise ValueError("1 <= ne <= 8")
        self.h = integrator
        self.H = tuple(float(v) for v in H_diag)
        self.ne = int(ne)
        self.fd_step = float(fd_step)

    # -- one evaluation of g(qe): 4 kernel launches
    def residual(self, qe, F_tip, M_tip, K0=None, work=None):
        h = self.h
        K = h.strain_from_modes(qe, out=None if work is None else work["K"])
        out = h.integrate_all(K, F_tip, M_tip, Q=None if work is None else work["Q"], m=None if work is None else work["m"],
                              want=("Q", "m"))
        rho = h.shape_residual(K, self.H, out["Q"], out["m"], M_tip, K0=K0, rho=None if work is None else work["rho"])
        g = h.project_onto_modes(rho, self.ne, out=None if work is None else work["g"])
        return g

    def solve(self, F_tip, M_tip, qe0=None, K0=None, tol: float = 1e-10, max_iter: int = 30, group=None):
        """Newton iteration until the GLOBAL rms of g over all ranks' rods is below tol.  Returns (qe, NewtonReport)."""
        h, ne = self.h, self.ne
        B = F_tip.shape[0]
        n = 3 * ne
        dev, f64 = F_tip.device, torch.float64
        qe = torch.zeros((B, n), dtype=f64, device=dev) if qe0 is None else qe0.clone()
        N, M = h.N, h.M
        work = {"K": torch.empty((B, 3, N), dtype=f64, device=dev), "Q": torch.empty((B, 4, M), dtype=f64, device=dev),
                "m": torch.empty((B, 3, M), dtype=f64, device=dev), "rho": torch.empty((B, 3, N), dtype=f64, device=dev),
                "g": torch.empty((B, n), dtype=f64, device=dev)}
        g0 = torch.empty((B, n), dtype=f64, device=dev)
        J = torch.empty((B, n, n), dtype=f64, device=dev)
        qp = torch.empty_like(qe)
        delta = torch.empty_like(qe)
        red = torch.zeros(2, dtype=f64, device=dev)
        count = torch.tensor([float(B * n)], dtype=f64, device=dev)
        if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
            torch.distributed.all_reduce(count, group=group)
        total_dof = float(count.item())
        rep = NewtonReport(iterations=0, converged=False)
        for it in range(max_iter + 1):
            g0.copy_(self.residual(qe, F_tip, M_tip, K0, work))
            rep.integrations += 1
            red[0] = (g0 * g0).sum()
            red[1] = g0.abs().max() if B else 0.0
            allreduce_residual(red, group)                     # the only collective: 16 bytes per iteration
            rms = float((red[0] / total_dof).sqrt().item())
            rep.rms_history.append(rms)
            rep.max_history.append(float(red[1].item()))
            if rms < tol:
                rep.converged = True
                break
            if it == max_iter:
                break
            for d in range(n):                                  # forward-difference Jacobian, column d
                qp.copy_(qe)
                qp[:, d] += self.fd_step
                gd = self.residual(qp, F_tip, M_tip, K0, work)
                J[:, :, d] = (gd - g0) / self.fd_step
                rep.integrations += 1
            h.solve_small_batched(J, g0, out=delta)             # per-rod n x n Newton system
            qe -= delta
            rep.iterations += 1
        return qe, rep
