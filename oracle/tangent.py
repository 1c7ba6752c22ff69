"""Exact tangent of the DISCRETE four-stage map and of the Galerkin residual of the static shape problem, in numpy.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): nothing in the product imports this.  Two restatements of the
Newton Jacobian of SURVEY 8 f1: (1) the exact tangent of the discrete stages -- the 3*ne tangent systems share the stage-1
operator A_NN(K) (main.cpp:55-88) -- pinned against central differences of the oracle's own residual; (2) the solve-free
form that sri_shape_jacobian computes on the GPU (jacobian_by_quadrature, at the end of this file), pinned against (1).

With Qs the stage-1 solution (stack [c*M+i], main.cpp:80-81) and a strain direction dK [3][N]:
    A_NN(K) dQs = 1/2 calA(dK) Qs                      (calA: the block-diagonal of updateA, main.cpp:72-75, linear in K)
    db_i        = dR(Q_i; dQ_i) Gamma_i               (R: Eigen's un-normalised toRotationMatrix, main.cpp:136; quadratic in Q)
    dm          = D_TT^-1 ( -(db x n) )               (rod_modeling.pdf eq. 1.18; n does not depend on K)
    drho_i      = H dK_i - dR_i^T m_i - R_i^T dm_i    (eq. 1.25), node 0: m = M_tip (dm = 0), base node: Q = q0 (dQ = 0)
    dg          = sum_i w_i Phi(x_i)^T drho_i         (eqs. 2.14 / 2.20, Clenshaw-Curtis weights)
"""
from __future__ import annotations

import numpy as np
from numpy.polynomial import legendre as _L


def rotation(q: np.ndarray) -> np.ndarray:
    """Eigen's Quaternion::toRotationMatrix without normalisation; q = (w, x, y, z) [..., 4] -> [..., 3, 3]."""
    w, x, y, z = np.moveaxis(q, -1, 0)
    R = np.empty(q.shape[:-1] + (3, 3))
    R[..., 0, 0] = 1 - 2 * (y * y + z * z); R[..., 0, 1] = 2 * (x * y - w * z); R[..., 0, 2] = 2 * (x * z + w * y)
    R[..., 1, 0] = 2 * (x * y + w * z); R[..., 1, 1] = 1 - 2 * (x * x + z * z); R[..., 1, 2] = 2 * (y * z - w * x)
    R[..., 2, 0] = 2 * (x * z - w * y); R[..., 2, 1] = 2 * (y * z + w * x); R[..., 2, 2] = 1 - 2 * (x * x + y * y)
    return R


def rotation_tangent(q: np.ndarray, dq: np.ndarray) -> np.ndarray:
    """dR(q; dq): R is I + a quadratic form in q, so the polarisation identity is exact."""
    return rotation(q + dq) - rotation(q) - rotation(dq) + np.eye(3)


def cc_weights(N: int) -> np.ndarray:
    n = N - 1
    w = np.zeros(N)
    for j in range(N):
        s = sum((1.0 if (k == 0 or 2 * k == n) else 2.0) / (1 - 4 * k * k) * np.cos(2 * k * j * np.pi / n) for k in range(n // 2 + 1))
        w[j] = 0.5 * (1.0 if j in (0, n) else 2.0) / n * s
    return w


def galerkin_residual_and_jacobian(oracle, qe, F_tip, M_tip, H_diag, ne: int, K0=None):
    """g [B][3 ne] and its exact Jacobian J [B][3 ne][3 ne] (J[b][:, d] = dg/dqe_d) for the oracle's discrete stages
    (Gamma = e1, q0 = identity, no distributed loads: the configuration of BASELINE configs[4])."""
    N, M, n = oracle.N, oracle.M, 3 * ne
    qe = np.ascontiguousarray(qe, dtype=np.float64).reshape(-1, n)
    B = qe.shape[0]
    H = np.asarray(H_diag, dtype=np.float64)
    x = oracle.chebyshev_points()
    P = np.stack([_L.legval(2 * x - 1, [0] * k + [1]) for k in range(ne)])            # [ne][N]
    w = cc_weights(N)
    S_T = np.linalg.inv(oracle.dn()[1:, 1:])                                          # D_TT^-1
    K = oracle.strain_from_modes(qe, ne)
    out = oracle.integrate_all(K, F_tip, M_tip, explicit_inverse=False, want=("Q", "n", "m"))
    A0 = oracle.assemble_A(np.zeros((3, N)))
    g = np.empty((B, n)); J = np.empty((B, n, n))
    e1 = np.array([1.0, 0.0, 0.0])
    for b in range(B):
        Qs = out["Q"][b].reshape(-1)                                                   # [c*M+i]
        A = oracle.assemble_A(K[b])
        # nodal quantities at all N nodes: Q (base node = q0), m (node 0 = M_tip)
        Qn = np.concatenate([out["Q"][b].T, [[1.0, 0.0, 0.0, 0.0]]])                   # [N][4]
        mn = np.concatenate([[M_tip[b]], out["m"][b].T])                               # [N][3]
        Rn = rotation(Qn)
        Kd = K[b] - (0.0 if K0 is None else K0[b])
        rho = H[:, None] * Kd - np.einsum("ikc,ik->ci", Rn, mn)                        # (R^T m)_c = sum_k R[k][c] m_k
        g[b] = np.einsum("ci,ki,i->ck", rho, P, w).reshape(n)
        for d in range(n):
            c, k = divmod(d, ne)
            dK = np.zeros((3, N)); dK[c] = P[k]
            v = (A0 - oracle.assemble_A(dK)) @ Qs                                      # 1/2 calA(dK) Qs
            dQs = np.linalg.solve(A, v)
            dQn = np.concatenate([dQs.reshape(4, M).T, np.zeros((1, 4))])              # [N][4], base node fixed
            dRn = rotation_tangent(Qn, dQn)
            db = dRn @ e1                                                              # [N][3]
            dm_inner = (S_T @ (-np.cross(db[1:], out["n"][b].T))).T                    # [3][M], nodes 1..N-1
            dmn = np.concatenate([np.zeros((1, 3)), dm_inner.T])                       # node 0: M_tip, no variation
            drho = H[:, None] * dK - np.einsum("ikc,ik->ci", dRn, mn) - np.einsum("ikc,ik->ci", Rn, dmn)
            J[b, :, d] = np.einsum("ci,ki,i->ck", drho, P, w).reshape(n)
    return g, J


def jacobian_by_quadrature(oracle, qe, F_tip, M_tip, H_diag, ne: int):
    """The same Jacobian WITHOUT any linear solve: the variation of the rotation is left-trivialised, dR = [dtheta]x R with
    dtheta' = R dK, dtheta(0) = 0, so every direction costs two contractions with the cached integration matrices
        dtheta = Dn_NN^-1 (R dK),     dm = D_TT^-1 ( -((dtheta x b) x n) ),     b = R Gamma,
        drho   = H dK - R^T (dm - dtheta x m).
    It is the tangent of the continuous problem collocated, not of the discrete map: the two agree to the discretisation
    error (1e-10 at N = 16, round-off at N = 32), far below what a Newton iteration can tell.  This is the formula of
    sri_shape_jacobian; returns J [B][3 ne][3 ne]."""
    N, M, n = oracle.N, oracle.M, 3 * ne
    qe = np.ascontiguousarray(qe, dtype=np.float64).reshape(-1, n)
    B = qe.shape[0]
    H = np.asarray(H_diag, dtype=np.float64)
    x = oracle.chebyshev_points()
    P = np.stack([_L.legval(2 * x - 1, [0] * k + [1]) for k in range(ne)])
    w = cc_weights(N)
    Dn = oracle.dn()
    S = np.linalg.inv(Dn[:M, :M]); S_T = np.linalg.inv(Dn[1:, 1:])
    K = oracle.strain_from_modes(qe, ne)
    out = oracle.integrate_all(K, F_tip, M_tip, explicit_inverse=False, want=("Q", "n", "m"))
    return jacobian_by_quadrature_from_state(out["Q"], out["n"], out["m"], M_tip, H, ne, P, w, S, S_T)


def jacobian_by_quadrature_from_state(Q, nn, mm, M_tip, H, ne, P, w, S, S_T, q0=None, Gamma=None):
    """q0 [B][4] (rotation of the base node, default identity) and Gamma [B][3][N] (default e1) as in sri_shape_jacobian."""
    B, _, M = Q.shape
    N, n = M + 1, 3 * ne
    e1 = np.array([1.0, 0.0, 0.0])
    J = np.empty((B, n, n))
    for b in range(B):
        qb = [1.0, 0.0, 0.0, 0.0] if q0 is None else q0[b]
        Rn = rotation(np.concatenate([Q[b].T, [qb]]))                                  # [N][3][3], base node = q0
        mn = np.concatenate([[M_tip[b]], mm[b].T])                                     # node 0: M_tip
        nj = nn[b].T                                                                   # nodes 1..N-1
        bn = Rn @ e1 if Gamma is None else np.einsum("ikc,ci->ik", Rn, Gamma[b])
        for d in range(n):
            c, k = divmod(d, ne)
            dK = np.zeros((3, N)); dK[c] = P[k]
            u = np.einsum("ikc,ci->ik", Rn, dK)
            th = np.concatenate([S @ u[:M], np.zeros((1, 3))])
            v = -np.cross(np.cross(th, bn)[1:], nj)
            dm = np.concatenate([np.zeros((1, 3)), S_T @ v])
            tv = dm - np.cross(th, mn)
            drho = H[:, None] * dK - np.einsum("ikc,ik->ci", Rn, tv)
            J[b, :, d] = np.einsum("ci,ki,i->ck", drho, P, w).reshape(n)
    return J
