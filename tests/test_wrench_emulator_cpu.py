"""The row-exact numpy model of the register-resident Gauss-Jordan kernels of the local-frame statics
(tools/wrench_gj_emulator.py: csrc/sri_wrench_gj_multi.cuh and csrc/sri_wrench_gj_static.cuh) against the oracle: pins on the
CPU the closed form of the preconditioned operator, both right-hand sides, the pivot-row trick, where unknown k ends up under
implicit pivoting, and what the growth check lets through."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
import wrench_gj_emulator as emu  # noqa: E402

TOL = 1e-12


def _inputs(o, N, B, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    x = o.chebyshev_points()
    K, F, Mt, fb = o.generate_rods(0x5EED, 100 + N, B)
    K = K * scale
    fbar = fb + 0.3 * rng.normal(size=(B, 3, 1)) * np.sin(2 * x)[None, None, :]
    lbar = 0.2 * rng.normal(size=(B, 3, 1)) * np.cos(x)[None, None, :]
    Gamma = np.stack([1 + 0.05 * np.sin(x), 0.03 * x, 0.02 * np.cos(x)])[None].repeat(B, axis=0)
    q0 = rng.normal(size=(B, 4)); q0 /= np.linalg.norm(q0, axis=1, keepdims=True)
    Q = rng.normal(size=(B, 4, N - 1)); Q /= np.linalg.norm(Q, axis=1, keepdims=True)
    return K, Q, F, Mt, q0, Gamma, fbar, lbar


def _rel(a, ref):
    return np.abs(a - ref).max() / np.abs(ref).max()


@pytest.mark.parametrize("static", [False, True])
@pytest.mark.parametrize("N", [9, 16, 17, 26])
def test_emulated_gauss_jordan_matches_oracle(make_oracle, N, static):
    o = make_oracle(N)
    D_TT, D_TI, S = o.operator(4), o.operator(5), o.operator(6)
    B = 3
    K, Q, F, Mt, q0, Gamma, fbar, lbar = _inputs(o, N, B, 7 * N)
    ref = o.wrench_local_solve(K, Q, F, Mt, q0=q0, Gamma=Gamma, fbar=fbar, lbar=lbar)
    ref0 = o.wrench_local_solve(K, Q, F, Mt)
    for b in range(B):
        lam, flag = emu.solve_rod(D_TT, D_TI, S, K[b], Q[b], F[b], Mt[b], q0=q0[b], Gamma=Gamma[b], fbar=fbar[b], lbar=lbar[b], static=static)
        assert not flag
        assert _rel(lam, ref[b]) <= TOL
        lam0, flag0 = emu.solve_rod(D_TT, D_TI, S, K[b], Q[b], F[b], Mt[b], static=static)   # every optional input absent
        assert not flag0 and _rel(lam0, ref0[b]) <= TOL


def test_preconditioned_operator_closed_form(make_oracle):
    o = make_oracle(12)
    D_TT, S = o.operator(4), o.operator(6)
    K = o.generate_rods(0x5EED, 3, 1)[0][0] * 7.0
    P = emu.preconditioned_operator(S, K)
    assert np.abs(P - np.kron(S, np.eye(3)) @ emu.raw_operator(D_TT, K)).max() <= 1e-12 * np.abs(P).max()


def test_static_order_growth_check(make_oracle):
    """Benchmark strains pass with a wide margin; very large curvatures are handed back (flag), and the row-pivoting form
    still solves them."""
    N = 20
    o = make_oracle(N)
    D_TT, D_TI, S = o.operator(4), o.operator(5), o.operator(6)
    K, Q, F, Mt, q0, Gamma, fbar, lbar = _inputs(o, N, 2, 5)
    P = emu.preconditioned_operator(S, K[0])
    _, L, _, flag = emu.eliminate(P, np.ones(3 * (N - 1)), static=True)
    off = L - np.diag(np.diag(L))
    assert not flag and np.abs(off).max() < 0.5
    Kbig = K * 80.0
    lam, flag = emu.solve_rod(D_TT, D_TI, S, Kbig[0], Q[0], F[0], Mt[0], fbar=fbar[0], static=True)
    assert flag
    lam_p, flag_p = emu.solve_rod(D_TT, D_TI, S, Kbig[0], Q[0], F[0], Mt[0], fbar=fbar[0], static=False)
    ref = o.wrench_local_solve(Kbig[:1], Q[:1], F[:1], Mt[:1], fbar=fbar[:1])[0]
    assert not flag_p and _rel(lam_p, ref) <= 1e-9


def test_implicit_pivoting_on_the_unpreconditioned_operator_moves_rows(make_oracle):
    """K = 0: the raw operator is D_TT (x) I3, whose pivots come from far below the diagonal; unknown k must be read from
    the row that was pivot at step k."""
    N = 10
    o = make_oracle(N)
    D_TT, D_TI, S = o.operator(4), o.operator(5), o.operator(6)
    K, Q, F, Mt, q0, Gamma, fbar, lbar = _inputs(o, N, 1, 11, scale=0.0)
    A = emu.raw_operator(D_TT, K[0])
    _, _, owner, flag = emu.eliminate(A, np.ones(3 * (N - 1)), static=False)
    assert not flag and (owner != np.arange(3 * (N - 1))).any() and sorted(owner) == list(range(3 * (N - 1)))
    lam, _ = emu.solve_rod(D_TT, D_TI, S, K[0], Q[0], F[0], Mt[0], fbar=fbar[0], lbar=lbar[0])
    ref = o.wrench_local_solve(K, Q, F, Mt, fbar=fbar, lbar=lbar)[0]
    assert _rel(lam, ref) <= TOL
