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


def test_newton_batched_jacobian_equals_column_by_column(h16, torch_mod):
    """jacobian="batched" integrates the 3 ne perturbed copies of the batch in one call of the hot path; every rod goes
    through the same kernels as in the column-by-column loop, so the iterates must be identical (with and without K0)."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200.newton import StaticShapeSolver
    B, ne = 333, 3
    rng = np.random.default_rng(11)
    F = np.zeros((B, 3)); F[:, 2] = -rng.uniform(0.1, 2.0, size=B)
    Mt = rng.uniform(-0.1, 0.1, size=(B, 3))
    K0 = 0.2 * rng.normal(size=(B, 3, 1)) * np.ones((1, 1, 16))
    tF, tM, tK0 = (torch_mod.from_numpy(np.ascontiguousarray(a)).cuda() for a in (F, Mt, K0))
    for k0 in (None, tK0):
        for use_graph in (False, True):
            cols = StaticShapeSolver(h16, (1.0, 1.0, 0.77), ne=ne, jacobian="columns")
            wide = StaticShapeSolver(h16, (1.0, 1.0, 0.77), ne=ne, jacobian="batched")
            q_c, rep_c = cols.solve(tF, tM, K0=k0, use_graph=use_graph)
            q_w, rep_w = wide.solve(tF, tM, K0=k0, use_graph=use_graph)
            assert rep_c.converged and rep_w.converged and rep_w.iterations == rep_c.iterations
            assert rep_w.integrations == rep_c.integrations
            assert torch_mod.equal(q_c, q_w)


@pytest.mark.parametrize("N", [16, 9, 32])
def test_galerkin_residual_equals_residual_then_projection(sri_lib, make_oracle, torch_mod, N):
    """sri_galerkin_residual = sri_project_onto_modes(sri_shape_residual(...)) without the nodal residual in memory: against
    the oracle's residual projected with numpy, against the two-kernel composition on the GPU, and its norms (fixed summation
    order: bitwise reproducible)."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
    o = make_oracle(N)
    rng = np.random.default_rng(60 + N)
    B = 1237
    K, F, Mt, fb = o.generate_rods(0x5EED, 77, B)
    ref = o.integrate_all(K, F, Mt, fbar=fb)
    K0 = 0.3 * rng.normal(size=(B, 3, N))
    q0 = None
    H = np.array([1.0, 0.9, 0.77])
    x = o.chebyshev_points()
    w = cc_weights(N)
    t = lambda a: torch_mod.from_numpy(np.ascontiguousarray(a)).cuda()
    with SpectralRodIntegrator(N, 0) as h:
        for ne, k0 in ((3, K0), (1, None), (8, K0)):
            rho_ref = o.shape_residual(K, H, ref["Q"], ref["m"], Mt, K0=k0)
            g_ref = np.einsum("bci,ki,i->bck", rho_ref, legendre_table(ne, 2 * x - 1), w).reshape(B, 3 * ne)
            red = torch_mod.full((2,), -1.0, dtype=torch_mod.float64, device="cuda")
            tk0 = None if k0 is None else t(k0)
            g = h.galerkin_residual(t(K), H, t(ref["Q"]), t(ref["m"]), t(Mt), ne, K0=tk0, reduce=red)
            g2 = h.project_onto_modes(h.shape_residual(t(K), H, t(ref["Q"]), t(ref["m"]), t(Mt), K0=tk0), ne)
            red_again = torch_mod.zeros(2, dtype=torch_mod.float64, device="cuda")
            g_again = h.galerkin_residual(t(K), H, t(ref["Q"]), t(ref["m"]), t(Mt), ne, K0=tk0, reduce=red_again)
            h.synchronize()
            g, g2, red = g.cpu().numpy(), g2.cpu().numpy(), red.cpu().numpy()
            scale = np.abs(g_ref).max()
            assert np.abs(g - g_ref).max() <= 1e-13 * scale
            assert np.abs(g - g2).max() <= 1e-14 * scale
            assert abs(red[0] - (g ** 2).sum()) <= 1e-13 * (g ** 2).sum() and red[1] == np.abs(g).max()
            assert torch_mod.equal(g_again.cpu(), torch_mod.from_numpy(g)) and (red_again.cpu().numpy() == red).all()
        # host buffers
        g_host = h.galerkin_residual(K[:7], H, ref["Q"][:7], ref["m"][:7], Mt[:7], 3)
        h.synchronize()
    rho_ref = o.shape_residual(K[:7], H, ref["Q"][:7], ref["m"][:7], Mt[:7])
    g_ref = np.einsum("bci,ki,i->bck", rho_ref, legendre_table(3, 2 * x - 1), w).reshape(7, 9)
    assert np.abs(g_host - g_ref).max() <= 1e-13 * np.abs(g_ref).max()


def test_generalised_forces_close_the_static_balance(h16, oracle16, torch_mod):
    """rod_modeling.pdf eq. 2.20: the Galerkin residual is the stiffness term plus the generalised internal forces,
    g = int Phi^T H (K - K0) dX + Q_ad, with Q_ad = -int Phi^T (couple part of the local-frame wrench)."""
    B, ne = 211, 4
    rng = np.random.default_rng(70)
    K, F, Mt, fb = oracle16.generate_rods(0x5EED, 500, B)
    ref = oracle16.integrate_all(K, F, Mt, fbar=fb)
    K0 = 0.1 * rng.normal(size=(B, 3, 16))
    H = np.array([1.0, 0.9, 0.77])
    x = oracle16.chebyshev_points()
    t = lambda a: torch_mod.from_numpy(np.ascontiguousarray(a)).cuda()
    lam = h16.wrench_local(t(ref["Q"]), t(ref["n"]), t(ref["m"]), t(F), t(Mt))
    qad = h16.generalised_forces(lam, ne)
    g = h16.galerkin_residual(t(K), H, t(ref["Q"]), t(ref["m"]), t(Mt), ne, K0=t(K0))
    stiff = h16.project_onto_modes(t(H[None, :, None] * (K - K0)), ne)
    h16.synchronize()
    lam_np = lam.cpu().numpy()
    ref_qad = -np.einsum("bci,ki,i->bck", lam_np[:, :3, :], legendre_table(ne, 2 * x - 1), cc_weights(16)).reshape(B, 3 * ne)
    assert np.abs(qad.cpu().numpy() - ref_qad).max() <= 1e-14 * np.abs(ref_qad).max()
    total = (stiff + qad).cpu().numpy()
    assert np.abs(total - g.cpu().numpy()).max() <= 1e-13 * np.abs(total).max()
    # host buffers
    qad_host = h16.generalised_forces(lam_np[:9], ne)
    h16.synchronize()
    assert np.abs(qad_host - ref_qad[:9]).max() <= 1e-14 * np.abs(ref_qad).max()


def test_native_newton_driver_matches_the_python_driver(h16, torch_mod):
    """sri_newton_static_shape (the loop inside the C ABI) takes the same steps as newton.StaticShapeSolver: same iteration
    count, same residual history and solution to round-off; device and host buffers; K0; the reduction callback."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200.newton import StaticShapeSolver
    B, ne = 417, 3
    rng = np.random.default_rng(21)
    F = np.zeros((B, 3)); F[:, 2] = -rng.uniform(0.1, 2.0, size=B)
    Mt = rng.uniform(-0.1, 0.1, size=(B, 3))
    K0 = 0.2 * rng.normal(size=(B, 3, 1)) * np.ones((1, 1, 16))
    tF, tM, tK0 = (torch_mod.from_numpy(np.ascontiguousarray(a)).cuda() for a in (F, Mt, K0))
    H = (1.0, 1.0, 0.77)
    for k0_t, k0_h in ((None, None), (tK0, K0)):
        q_py, rep_py = StaticShapeSolver(h16, H, ne=ne).solve(tF, tM, K0=k0_t, use_graph=False)
        q_c, rep_c = h16.newton_static_shape(tF, tM, ne, H, K0=k0_t)
        h16.synchronize()
        assert rep_c["converged"] and rep_c["iterations"] == rep_py.iterations and rep_c["integrations"] == rep_py.integrations
        # (torch divides by the step through a reciprocal: the forward differences agree to ~1e-10, not to the bit)
        big = [i for i, v in enumerate(rep_py.rms_history) if v > 1e-6]
        assert np.allclose(np.array(rep_c["rms_history"])[big], np.array(rep_py.rms_history)[big], rtol=1e-6)
        assert np.abs(q_c.cpu().numpy() - q_py.cpu().numpy()).max() <= 1e-11
        q_h, rep_h = h16.newton_static_shape(F, Mt, ne, H, K0=k0_h)          # host buffers
        assert rep_h["iterations"] == rep_c["iterations"] and np.abs(q_h - q_c.cpu().numpy()).max() <= 1e-15
    # the reduction callback sees this rank's norms and may replace them (two identical ranks: sum doubles, max stays)
    seen = []
    def two_ranks(norms):
        seen.append(norms.copy()); norms[0] *= 2.0
    q_r, rep_r = h16.newton_static_shape(tF, tM, ne, H, total_dof=2 * B * 3 * ne, allreduce=two_ranks)
    assert rep_r["converged"] and len(seen) == rep_r["iterations"] + 1
    q_1, rep_1 = h16.newton_static_shape(tF, tM, ne, H)
    assert np.allclose(rep_r["rms_history"], rep_1["rms_history"], rtol=1e-12) and torch_mod.equal(q_r, q_1)
    # an empty shard still takes part in every reduction
    calls = []
    empty = torch_mod.zeros((0, 3), dtype=torch_mod.float64, device="cuda")
    _, rep_e = h16.newton_static_shape(empty, empty, ne, H, total_dof=10, max_iter=3, allreduce=lambda nrm: (calls.append(1), nrm.__setitem__(0, 1.0)))
    assert not rep_e["converged"] and rep_e["iterations"] == 3 and len(calls) == 4


def test_cpp_host_drives_the_newton_solve(sri_lib):
    """examples/newton_main.cpp: C++ host code -> sri_newton_static_shape; pure tip moments must give K = H^-1 M_tip."""
    import subprocess
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    exe = root / "examples" / "newton_main_gpu"
    if not exe.exists():
        import __graft_entry__ as g
        g._build_cpp_example(root / "experimental_gpu_programming_for_a_spectral_numerical_integration_b200" / "libsri_cuda.so")
    res = subprocess.run([str(exe)], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert res.stdout.startswith("converged 1")


@pytest.mark.parametrize("N,ne", [(16, 3), (16, 4), (9, 2), (32, 3), (16, 1), (16, 8), (12, 5), (5, 2), (16, 6), (13, 7)])
def test_shape_jacobian_matches_the_restated_quadrature_formula(sri_lib, make_oracle, torch_mod, N, ne):
    """sri_shape_jacobian (solve-free analytic Jacobian) against oracle/tangent.py's restatement on the same stage outputs;
    that restatement is pinned on the CPU against the exact tangent of the discrete map and against central differences."""
    from numpy.polynomial import legendre as L
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
    from oracle.tangent import cc_weights as ccw, jacobian_by_quadrature_from_state
    o = make_oracle(N)
    M = N - 1
    B = 37
    rng = np.random.default_rng(80 + N + ne)
    qe = 0.7 * rng.normal(size=(B, 3 * ne))
    F = rng.uniform(-1, 1, size=(B, 3)); Mt = rng.uniform(-0.5, 0.5, size=(B, 3))
    H = np.array([1.0, 0.9, 0.77])
    K = o.strain_from_modes(qe, ne)
    ref = o.integrate_all(K, F, Mt, explicit_inverse=False, want=("Q", "n", "m"))
    x = o.chebyshev_points()
    P = np.stack([L.legval(2 * x - 1, [0] * k + [1]) for k in range(ne)])
    Dn = o.dn()
    J_ref = jacobian_by_quadrature_from_state(ref["Q"], ref["n"], ref["m"], Mt, H, ne, P, ccw(N), np.linalg.inv(Dn[:M, :M]),
                                              np.linalg.inv(Dn[1:, 1:]))
    t = lambda a: torch_mod.from_numpy(np.ascontiguousarray(a)).cuda()
    with SpectralRodIntegrator(N, 0) as h:
        J = h.shape_jacobian(t(ref["Q"]), t(ref["n"]), t(ref["m"]), t(Mt), ne, H)
        J_host = h.shape_jacobian(ref["Q"][:5], ref["n"][:5], ref["m"][:5], Mt[:5], ne, H)
        h.synchronize()
        # optional inputs: rotation of the base node and a nodal Gamma (shearable rod)
        q0 = rng.normal(size=(B, 4)); q0 /= np.linalg.norm(q0, axis=1, keepdims=True)
        Gam = np.tile(np.array([1.0, 0.04, -0.03])[None, :, None], (B, 1, N)) + 0.01 * rng.normal(size=(B, 3, N))
        ref2 = o.integrate_all(K, F, Mt, q0=q0, Gamma=Gam, explicit_inverse=False, want=("Q", "n", "m"))
        J2 = h.shape_jacobian(t(ref2["Q"]), t(ref2["n"]), t(ref2["m"]), t(Mt), ne, H, q0=t(q0), Gamma=t(Gam))
        h.synchronize()
    scale = np.abs(J_ref).max()
    assert np.abs(J.cpu().numpy() - J_ref).max() <= 1e-12 * scale
    assert np.abs(J_host - J_ref[:5]).max() <= 1e-12 * scale
    J2_ref = jacobian_by_quadrature_from_state(ref2["Q"], ref2["n"], ref2["m"], Mt, H, ne, P, ccw(N), np.linalg.inv(Dn[:M, :M]),
                                               np.linalg.inv(Dn[1:, 1:]), q0=q0, Gamma=Gam)
    assert np.abs(J2.cpu().numpy() - J2_ref).max() <= 1e-12 * np.abs(J2_ref).max()


def test_newton_with_the_analytic_jacobian(h16, torch_mod):
    """jacobian="analytic" (one integration per iteration) reaches the same shapes as the finite-difference Newton, in both
    drivers, in at most one more iteration."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200.newton import StaticShapeSolver
    B, ne = 600, 3
    rng = np.random.default_rng(31)
    F = np.zeros((B, 3)); F[:, 2] = -rng.uniform(0.1, 2.0, size=B); F[:, 1] = rng.uniform(-0.3, 0.3, size=B)
    Mt = rng.uniform(-0.2, 0.2, size=(B, 3))
    K0 = 0.2 * rng.normal(size=(B, 3, 1)) * np.ones((1, 1, 16))
    tF, tM, tK0 = (torch_mod.from_numpy(np.ascontiguousarray(a)).cuda() for a in (F, Mt, K0))
    H = (1.0, 1.0, 0.77)
    for k0 in (None, tK0):
        q_fd, rep_fd = StaticShapeSolver(h16, H, ne=ne).solve(tF, tM, K0=k0)
        for use_graph in (False, True):
            q_an, rep_an = StaticShapeSolver(h16, H, ne=ne, jacobian="analytic").solve(tF, tM, K0=k0, use_graph=use_graph)
            assert rep_an.converged and rep_an.iterations <= rep_fd.iterations + 1
            assert rep_an.integrations == rep_an.iterations + 1
            assert np.abs(q_an.cpu().numpy() - q_fd.cpu().numpy()).max() <= 1e-9
        q_c, rep_c = h16.newton_static_shape(tF, tM, ne, H, K0=k0, fd_step=0.0)
        assert rep_c["converged"] and rep_c["iterations"] == rep_an.iterations and rep_c["integrations"] == rep_c["iterations"] + 1
        assert np.abs(q_c.cpu().numpy() - q_fd.cpu().numpy()).max() <= 1e-9
