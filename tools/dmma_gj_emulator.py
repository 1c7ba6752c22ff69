"""Lane-exact numpy emulation of the DMMA Gauss-Jordan of csrc/sri_fused16_dmma.cuh (stage 1, one rod per warp).

Purpose: the CUDA kernel's index algebra (which lane holds which entry, every shuffle source, every sign) is
checked HERE, on the CPU, against the oracle, before any GPU time is spent.  Registers are arrays over the 32
lanes; `shfl` and `mma_m8n8k4` follow the PTX fragment layouts
    A (8x4, row):  lane l holds A[l>>2][l&3]
    B (4x8, col):  lane l holds B[l&3][l>>2]
    C (8x8):       lane l holds C[l>>2][2*(l&3) + e], e = 0, 1.

Real layout of the M x M quaternion system  sum_j Q_j (x) c_ij = b_i  (DESIGN.md section 1):
    Cr[4*i + r][j] = component r of c_ij,  j = 0..14;  Cr[4*i + r][15] = component r of b_i
16 tiles of 8 x 8: tile (t, ct) = rows 8t..8t+7 (quaternion rows 2t, 2t+1), columns 8ct..8ct+7.

Step k (static pivot order, growth check instead of a search):
    u'_j  = u_j (x) c_kk^-1                 (pivot row normalised: one DMMA per column tile)
    c_ij -= u'_j (x) c_ik   for all rows i  (vec(a (x) b) = Rmat(b) vec(a): [Rmat(c_ik) stacked] * [-U'] : DMMAs)
    pivot row itself: Rmat(c_kk - 1), which leaves exactly u'_j.
"""
from __future__ import annotations

import numpy as np

LANES = np.arange(32)
RHO, CP = LANES >> 2, LANES & 3
HI, R = RHO >> 2, RHO & 3

# Rmat(b)[r][s] = SG[r][s] * b[r ^ s]   (vec(a (x) b) = Rmat(b) vec(a), components w, x, y, z)
SG = np.array([[1, -1, -1, -1], [1, 1, 1, -1], [1, -1, 1, 1], [1, 1, -1, 1]], dtype=np.float64)


def shfl(v, src):
    return v[src]


def mma_m8n8k4(c0, c1, a, b):
    A = np.zeros((8, 4)); B = np.zeros((4, 8)); C = np.zeros((8, 8))
    A[RHO, CP] = a
    B[CP, RHO] = b
    C[RHO, 2 * CP] = c0
    C[RHO, 2 * CP + 1] = c1
    D = C.copy()
    for q in range(4):  # sequential FMA chain over k, as the hardware accumulates
        D = D + np.outer(A[:, q], B[q, :])
    return D[RHO, 2 * CP], D[RHO, 2 * CP + 1]


def tables(S, g, M):
    """Stx[16][16]: -1/2 S_ij (i, j < M), column 15 = g_i, zero elsewhere."""
    Stx = np.zeros((16, 16))
    Stx[:M, :M] = -0.5 * S
    Stx[:M, 15] = g
    return Stx


def solve_rod(Stx, K, q0, M, growth=8.0):
    """K: [3][N]; returns (Q [M][4], flagged)."""
    kx = np.zeros((4, 16))
    kx[1:, :M] = K[:, :M]
    kx[:, 15] = q0
    # ---- assembly -------------------------------------------------------------------------------------------
    c = np.zeros((8, 2, 2, 32))
    for t in range(8):
        for ct in range(2):
            for e in range(2):
                i = 2 * t + HI
                j = 8 * ct + 2 * CP + e
                diag = ((R == 0) & (i == j) & (j < 15)).astype(np.float64)
                c[t, ct, e] = Stx[i, j] * kx[R, j] + diag
    flagged = False
    # lane constants
    sgL = SG[R, CP]
    srcL_base = 16 * (LANES >> 4) + 4 * (R ^ CP)
    n_even = (RHO & 1) == 0
    sp = RHO >> 1  # s' of the normalisation operand (even output columns only)
    idxN = sp ^ CP
    conj = np.where(idxN == 0, 1.0, -1.0)
    sgN = np.where(n_even, -SG[sp, CP] * conj, 0.0)
    for k in range(M):
        kt, kh, kc, kcp, ke = k >> 1, k & 1, k >> 3, (k & 7) >> 1, k & 1
        cts = [ct for ct in range(2) if 8 * ct + 7 > k]  # column tiles with live columns (j > k), rhs included
        # 1. pivot row -> B fragments
        src = 16 * kh + 4 * CP + (LANES >> 3)
        Ub = {}
        for ct in range(kc, 2):  # the pivot column's tile is needed for c_kk even when it has no live column left
            v0 = shfl(c[kt, ct, 0], src)
            v1 = shfl(c[kt, ct, 1], src)
            Ub[ct] = np.where((RHO & 1) == 1, v1, v0)
        # 2. this lane's entry of -Rmat(conj c_kk), straight from the pivot's tile; un0 = -U (x) conj(c_kk) by DMMA;
        #    its (column k, w) entry is -|c_kk|^2
        pcs = shfl(c[kt, kc, ke], 16 * kh + 4 * idxN + kcp)
        bn0 = sgN * pcs
        un0 = {ct: mma_m8n8k4(np.zeros(32), np.zeros(32), Ub[ct], bn0)[0] for ct in Ub}
        nu = -shfl(un0[kc], np.full(32, 4 * (k & 7)))
        nrm = nu
        inv = 1.0 / nu
        # 3. L fragments (column k of every row tile), before anything is updated
        srcL = srcL_base + kcp
        La = []
        colmax = np.zeros(32)
        for t in range(8):
            v = shfl(c[t, kc, ke], srcL)
            if t > kt:
                colmax = np.maximum(colmax, np.abs(v))
            elif t == kt:
                colmax = np.maximum(colmax, np.where(HI > kh, np.abs(v), 0.0))
            a = sgL * v
            if t == kt:
                a = a - ((HI == kh) & (R == CP)).astype(np.float64)
            La.append(a)
        if not (colmax.max() ** 2 <= growth * growth * nrm[0]) or not (nrm[0] > 1e-300):
            flagged = True
        # 4. normalise the pivot row (negated), 5. update
        for ct in cts:
            d0 = un0[ct] * inv
            for t in range(8):
                c[t, ct, 0], c[t, ct, 1] = mma_m8n8k4(c[t, ct, 0], c[t, ct, 1], La[t], d0)
    Q = np.zeros((16, 4))
    for t in range(8):
        sel = CP == 3
        Q[2 * t + HI[sel], R[sel]] = c[t, 1, 1][sel]
    return Q[:M], flagged


def main():
    import sys
    from pathlib import Path
    sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
    from oracle.oracle import Oracle

    rng = np.random.default_rng(7)
    for N in (16, 9, 4):
        M = N - 1
        orc = Oracle(N)
        S = orc.operator(3)
        Dn_IN = orc.operator(2)
        g = -S @ Dn_IN
        Stx = tables(S, g, M)
        B = 6
        x = orc.chebyshev_points()
        al, be = rng.uniform(-2, 2, (B, 3, 1)), rng.uniform(-2, 2, (B, 3, 1))
        K = al + be * (2 * x[None, None, :] - 1)
        q0 = rng.normal(size=(B, 4)); q0 /= np.linalg.norm(q0, axis=1, keepdims=True)
        ref = orc.integrate_all(K, q0=q0, want=("Q",))["Q"]
        worst = 0.0
        for b in range(B):
            Q, flagged = solve_rod(Stx, K[b], q0[b], M)
            err = np.abs(Q.T - ref[b]).max() / np.abs(ref[b]).max()
            worst = max(worst, err)
            assert not flagged
        print(f"N={N}: max rel err vs oracle {worst:.2e}")
        assert worst < 1e-12
    # large curvature: the growth check must fire for some rods
    orc = Oracle(16); S = orc.operator(3); g = -S @ orc.operator(2); Stx = tables(S, g, 15)
    x = orc.chebyshev_points()
    nflag = 0; worst = 0.0
    for b in range(20):
        K = rng.uniform(-70, 70, (3, 1)) + rng.uniform(-30, 30, (3, 1)) * (2 * x[None, :] - 1)
        q0 = np.array([1.0, 0, 0, 0])
        ref = orc.integrate_all(K[None], q0=q0[None], want=("Q",))["Q"][0]
        Q, flagged = solve_rod(Stx, K, q0, 15)
        nflag += flagged
        if not flagged:
            worst = max(worst, np.abs(Q.T - ref).max() / np.abs(ref).max())
    print(f"|K|<=70: {nflag}/20 rods flagged for the pivoting kernel; worst unflagged rel err {worst:.2e}")


if __name__ == "__main__":
    main()
