"""GPU tests of the static-shape Newton driver (BASELINE configs[4]) and its two helper kernels."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def cc_weights(N):
    n = N - 1
    w = np.zeros(N)
    for j in range(N):
        s = sum((1.0 if (k == 0 or 2 * k == n) else 2.0) / (1 - 4 * k * k) * np.cos(2 * k * j * np.pi / n) for k in range(n // 2 + 1))
        w[j] = 0.5 * (1.0 if j in (0, n) else 2.0) / n * s
    return w


def legendre_table(ne, t):
    from numpy.polynomial import legendre as L
    return np.stack([L.legval(t, [0] * k + [1]) for k in range(ne)])  # [ne][N]


@pytest.fixture(scope="module")
def torch_mod():
    import torch
    return torch


@pytest.fixture(scope="module")
def h16(sri_lib):
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
    h = SpectralRodIntegrator(16, 0)
    yield h
    h.close()


def test_projection_matches_numpy_and_is_exact_for_polynomials(h16, oracle16, torch_mod):
    rng = np.random.default_rng(0)
    x = oracle16.chebyshev_points()
    w = cc_weights(16)
    assert abs(w.sum() - 1) < 1e-15
    for ne in (1, 3, 4, 8):
        f = rng.normal(size=(33, 3, 16))
        P = legendre_table(ne, 2 * x - 1)
        ref = np.einsum("bci,ki,i->bck", f, P, w).reshape(33, 3 * ne)
        got = h16.project_onto_modes(torch_mod.from_numpy(f).cuda(), ne).cpu().numpy()
        assert np.abs(got - ref).max() < 1e-14
    # orthogonality of the Legendre modes on [0,1]: int P_k P_l = delta_kl / (2k+1)
    P = legendre_table(4, 2 * x - 1)
    f = np.broadcast_to(P[2], (1, 3, 16)).copy()
    got = h16.project_onto_modes(torch_mod.from_numpy(f).cuda(), 4).cpu().numpy().reshape(3, 4)
    assert np.abs(got - np.array([0, 0, 1 / 5, 0])).max() < 1e-14


def test_small_batched_solve_matches_lapack(h16, torch_mod):
    rng = np.random.default_rng(1)
    for n in (1, 3, 9, 12, 24):
        A = rng.normal(size=(257, n, n)) + 0.1 * np.eye(n)
        b = rng.normal(size=(257, n))
        info = torch_mod.full((257,), -1, dtype=torch_mod.int32, device="cuda")
        x = h16.solve_small_batched(torch_mod.from_numpy(A.copy()).cuda(), torch_mod.from_numpy(b).cuda(), info=info)
        ref = np.linalg.solve(A, b[..., None])[..., 0]
        assert (info.cpu().numpy() == 0).all()
        assert np.abs(x.cpu().numpy() - ref).max() <= 1e-9 * np.abs(ref).max()


def test_newton_pure_tip_moment_gives_circular_arc(h16, torch_mod):
    """F_tip = 0, f = l = 0  =>  m == M_tip and, for a moment in the isotropic bending plane, K == H^-1 M_tip."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200.newton import StaticShapeSolver
    B, ne = 64, 3
    rng = np.random.default_rng(2)
    Mt = np.zeros((B, 3)); Mt[:, :2] = rng.uniform(-1.5, 1.5, size=(B, 2))
    H = (1.3, 1.3, 0.77)
    solver = StaticShapeSolver(h16, H, ne=ne)
    qe, rep = solver.solve(torch_mod.zeros((B, 3), dtype=torch_mod.float64, device="cuda"), torch_mod.from_numpy(Mt).cuda())
    assert rep.converged and rep.iterations <= 6, rep
    want = np.zeros((B, 3, ne)); want[:, 0, 0] = Mt[:, 0] / H[0]; want[:, 1, 0] = Mt[:, 1] / H[1]
    assert np.abs(qe.cpu().numpy().reshape(B, 3, ne) - want).max() < 1e-9


def test_newton_tip_force_matches_cpu_replica(h16, oracle16, torch_mod):
    """Same Newton iteration restated with the oracle + numpy on the host, for a handful of tip-loaded rods."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200.newton import StaticShapeSolver
    B, ne, n = 6, 4, 12
    rng = np.random.default_rng(3)
    F = np.zeros((B, 3)); F[:, 2] = -rng.uniform(0.2, 2.0, size=B); F[:, 1] = rng.uniform(-0.5, 0.5, size=B)
    Mt = rng.uniform(-0.2, 0.2, size=(B, 3))
    H = np.array([1.0, 1.0, 0.77])
    solver = StaticShapeSolver(h16, H, ne=ne)
    qe, rep = solver.solve(torch_mod.from_numpy(F).cuda(), torch_mod.from_numpy(Mt).cuda(), tol=1e-11)
    assert rep.converged and rep.iterations <= 12, rep
    assert rep.integrations == (rep.iterations * (n + 1) + 1)

    x = oracle16.chebyshev_points(); w = cc_weights(16); P = legendre_table(ne, 2 * x - 1)

    def g_cpu(q):
        K = oracle16.strain_from_modes(q, ne)
        out = oracle16.integrate_all(K, F, Mt, explicit_inverse=False, want=("Q", "m", "n"))
        rho = oracle16.shape_residual(K, H, out["Q"], out["m"], Mt)
        return np.einsum("bci,ki,i->bck", rho, P, w).reshape(B, n)

    q = np.zeros((B, n))
    for _ in range(15):
        g0 = g_cpu(q)
        if np.sqrt((g0 ** 2).mean()) < 1e-11:
            break
        J = np.empty((B, n, n))
        for d in range(n):
            qp = q.copy(); qp[:, d] += 1e-6
            J[:, :, d] = (g_cpu(qp) - g0) / 1e-6
        q -= np.linalg.solve(J, g0[..., None])[..., 0]
    assert np.abs(qe.cpu().numpy() - q).max() < 1e-8
    # and the converged shape is a real equilibrium: the nodal residual is at discretisation level
    assert np.abs(g_cpu(qe.cpu().numpy())).max() < 1e-9


def test_newton_cuda_graph_replay_equals_eager_and_is_reused(h16, torch_mod):
    """The captured iteration (CUDA graph) must reproduce the eager iteration bit for bit, and a second solve with the same
    batch size must replay the cached graph from its first iteration on."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200.newton import StaticShapeSolver
    B, ne = 500, 3
    rng = np.random.default_rng(5)
    F = np.zeros((B, 3)); F[:, 2] = -rng.uniform(0.1, 2.0, size=B)
    Mt = rng.uniform(-0.1, 0.1, size=(B, 3))
    tF, tM = torch_mod.from_numpy(F).cuda(), torch_mod.from_numpy(Mt).cuda()
    eager = StaticShapeSolver(h16, (1.0, 1.0, 0.77), ne=ne)
    q_e, rep_e = eager.solve(tF, tM, use_graph=False)
    graphed = StaticShapeSolver(h16, (1.0, 1.0, 0.77), ne=ne)
    q_g, rep_g = graphed.solve(tF, tM)
    assert rep_e.converged and rep_g.converged and rep_g.iterations == rep_e.iterations >= 3
    assert torch_mod.equal(q_e, q_g)
    assert rep_g.rms_history == rep_e.rms_history
    ws = next(iter(graphed._cache.values()))
    assert ws["graph"] is not None
    # second solve, other loads, same shape: graph reused (no new capture), still correct
    g_before = ws["graph"]
    q2_g, rep2_g = graphed.solve(tF * 0.5, tM)
    q2_e, rep2_e = eager.solve(tF * 0.5, tM, use_graph=False)
    assert next(iter(graphed._cache.values()))["graph"] is g_before
    assert rep2_g.converged and torch_mod.equal(q2_e, q2_g)
    # the integrator handle is usable outside the graph afterwards
    K = torch_mod.zeros((4, 3, 16), dtype=torch_mod.float64, device="cuda")
    out = h16.integrate_all(K, tF[:4], tM[:4])
    h16.synchronize()
    assert np.isfinite(out["Q"].cpu().numpy()).all()
