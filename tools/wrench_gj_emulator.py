"""Row-exact numpy model of the register-resident Gauss-Jordan kernels of the local-frame statics solve
(csrc/sri_wrench_gj_multi.cuh: implicit partial pivoting; csrc/sri_wrench_gj_static.cuh: static order on the operator
preconditioned by D_TT^-1, growth check).  One Python "row" = one lane of the kernel: a sliding window whose slot 0 is the
current pivot column, the pivot row normalised by the same update as the others (multiplier 1 - 1/pivot), multipliers kept
by (step, row) and replayed on the second right-hand side.  Used by tests/test_wrench_emulator_cpu.py against the oracle:
it pins the algebra (closed form of the preconditioned operator, both right-hand sides, where unknown k ends up) on the CPU.

Conventions (sri.h): K, Gamma, fbar, lbar [3][N] by node, node 0 = the tip where F_tip, M_tip act; Q [4][M] = quaternions of
nodes 0..M-1, q0 = the last node; Lambda [6][N], couple first.  Operators column-major as sri_get_operator: D_TT [M][M] =
Dn[1:,1:], D_TI [M] = Dn[1:,0], S = D_TT^-1."""
import numpy as np


def rot(q):
    """Rotation matrix of a quaternion (w, x, y, z), as Eigen's toRotationMatrix."""
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def hat(k):
    return np.array([[0.0, -k[2], k[1]], [k[2], 0.0, -k[0]], [-k[1], k[0], 0.0]])


def raw_operator(D_TT, K):
    """D_TT (x) I3 + blockdiag(K^_j), j = nodes 1..M: what the row-pivoting kernels eliminate."""
    M = D_TT.shape[0]
    A = np.kron(D_TT, np.eye(3))
    for j in range(M):
        A[3 * j:3 * j + 3, 3 * j:3 * j + 3] += hat(K[:, j + 1])
    return A


def preconditioned_operator(S, K):
    """Closed form of (S (x) I3)(D_TT (x) I3 + blockdiag K^): block (i, j) = delta_ij I3 + S_ij K^_j."""
    M = S.shape[0]
    P = np.eye(3 * M)
    for i in range(M):
        for j in range(M):
            P[3 * i:3 * i + 3, 3 * j:3 * j + 3] += S[i, j] * hat(K[:, j + 1])
    return P


def eliminate(A, b, static, growth=4.0):
    """Rolled Gauss-Jordan on sliding windows.  Returns (x by unknown index, L [step][row], owner row of every step, flag).
    static=False: implicit partial pivoting (largest |a| among the rows not yet used, first row on ties), flag = singular.
    static=True: pivot row k at step k, flag = some multiplier exceeded `growth` (or was not finite)."""
    n = A.shape[0]
    win = A.copy()                 # win[r, j] = slot j of row r's window (slot 0 = column k at step k)
    rhs = b.astype(float).copy()
    used = np.zeros(n, bool)
    L = np.zeros((n, n))
    owner = np.zeros(n, int)
    flag = False
    for k in range(n):
        if static:
            p = k
        else:
            cand = np.where(used, -1.0, np.abs(win[:, 0]))
            p = int(np.argmax(cand))   # first maximum = lowest row
            if not (cand[p] > 0.0) or not np.isfinite(cand[p]):
                flag = True
        used[p] = True
        owner[k] = p
        with np.errstate(all="ignore"):
            inv = 1.0 / win[p, 0]
            ml = win[:, 0] * inv
        ml[p] = 1.0 - inv
        L[k] = ml
        if static:
            others = np.delete(ml, p)
            if not np.all(np.abs(others) <= growth):   # (NaN compares false)
                flag = True
        u, ub = win[p].copy(), rhs[p]
        win = win - np.outer(ml, u)    # row p: u - (1 - inv) u = u / pivot
        rhs = rhs - ml * ub
        win = np.concatenate([win[:, 1:], np.zeros((n, 1))], axis=1)   # slide: slot 0 is now column k + 1
    x = np.empty(n)
    x[np.arange(n)] = rhs[owner]       # the row that was pivot at step k holds unknown k
    return x, L, owner, flag


def replay(L, owner, b):
    """Second right-hand side through the stored multipliers: the pivot value of step k is the owner row's current entry."""
    b = b.astype(float).copy()
    for k in range(L.shape[0]):
        b = b - L[k] * b[owner[k]]
    return b[owner]


def solve_rod(D_TT, D_TI, S, K, Q, F_tip, M_tip, q0=None, Gamma=None, fbar=None, lbar=None, static=False, growth=4.0):
    """One rod as the kernels compute it.  Returns (Lambda [6][N], flag)."""
    M = D_TT.shape[0]
    N = M + 1
    R = [rot(Q[:, t]) for t in range(M)] + [rot(q0) if q0 is not None else np.eye(3)]
    N0, C0 = R[0].T @ F_tip, R[0].T @ M_tip
    g1 = np.zeros(3 * M)
    if fbar is not None:
        for i in range(M):
            g1[3 * i:3 * i + 3] = R[i + 1].T @ fbar[:, i + 1]
    if static:
        A = preconditioned_operator(S, K)
        SI = np.kron(S, np.eye(3))
        sdti = S @ D_TI
        b1 = -(SI @ g1) - np.kron(sdti, N0)
    else:
        A = raw_operator(D_TT, K)
        b1 = -g1 - np.kron(D_TI, N0)
    xN, L, owner, flag = eliminate(A, b1, static, growth)
    g2 = np.zeros(3 * M)
    for i in range(M):
        gam = Gamma[:, i + 1] if Gamma is not None else np.array([1.0, 0.0, 0.0])
        g2[3 * i:3 * i + 3] = np.cross(gam, xN[3 * i:3 * i + 3])
        if lbar is not None:
            g2[3 * i:3 * i + 3] += R[i + 1].T @ lbar[:, i + 1]
    b2 = (-(SI @ g2) - np.kron(sdti, C0)) if static else (-g2 - np.kron(D_TI, C0))
    xC = replay(L, owner, b2)
    lam = np.empty((6, N))
    lam[:3, 0], lam[3:, 0] = C0, N0
    lam[:3, 1:] = xC.reshape(M, 3).T
    lam[3:, 1:] = xN.reshape(M, 3).T
    return lam, flag
