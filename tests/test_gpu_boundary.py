"""GPU tests of the drop-in boundary beyond the four stages: updateA on the reference surface, error
reporting on host-buffer calls, handles that share one GPU, the multi-device host path, the NCCL hook of the Newton
driver, and the rod-length scaling of SURVEY 8 f3."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import rel_err

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "experimental_gpu_programming_for_a_spectral_numerical_integration_b200"
TOL = 1e-12


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


# ---- updateA (main.cpp:55-88) ------------------------------------------------------------------------------------------

def test_assemble_A_matches_reference_updateA_and_oracle(h16, oracle16, make_oracle, torch_mod):
    """sri_assemble_A against the reference's own updateA output (golden fixture, rods 0..3: bit-exact), against the oracle
    on a random batch (bit-exact: same formula, no rounding freedom), device and host buffers, and N = 32."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
    g = np.load(ROOT / "tests" / "golden" / "reference_random_rods.npz")
    A = h16.assemble_A(torch_mod.from_numpy(g["K"][:4].copy()).cuda()).cpu().numpy()
    assert np.array_equal(A, g["A_NN"])
    K = oracle16.generate_rods(0x5EED, 100, 37)[0]
    A_host = h16.assemble_A(K)  # numpy in, numpy out: staged by the library
    for b in range(K.shape[0]):
        assert np.array_equal(A_host[b], oracle16.assemble_A(K[b]))
    o32 = make_oracle(32)
    K32 = o32.generate_rods(0x5EED, 0, 5)[0]
    with SpectralRodIntegrator(32, 0) as h32:
        A32 = h32.assemble_A(torch_mod.from_numpy(K32).cuda()).cpu().numpy()
    for b in range(5):
        assert np.array_equal(A32[b], o32.assemble_A(K32[b]))
    # the assembled operator is the one the integration kernels solve: A_NN Q = -D_IN q0
    Q = h16.integrate_quaternions(torch_mod.from_numpy(K).cuda()).cpu().numpy().reshape(K.shape[0], -1)
    rhs = -np.kron(np.array([1.0, 0, 0, 0]), oracle16.operator(2))
    for b in range(K.shape[0]):
        res = A_host[b] @ Q[b] - rhs
        assert np.abs(res).max() <= 2e-12 * np.abs(A_host[b]).max()


def test_cpp_updateA_keeps_the_reference_call_shape(sri_lib, tmp_path):
    """include/sri_reference_api.hpp::updateA<N,ne>(qe, A_NN, D_NN): C++ host code builds D_NN as main.cpp:98 does, calls
    updateA as main.cpp:103 does and dumps A_NN; compared with the golden updateA output of the real reference."""
    import shutil
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    g = np.load(ROOT / "tests" / "golden" / "reference_random_rods.npz")
    qe = g["qe"][2]
    src = tmp_path / "update_a.cpp"
    src.write_text("""#include "sri_reference_api.hpp"
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv) {
    constexpr int N = 16, M = N - 1, n = 4 * M;
    std::array<double, 9> qe;
    for (int i = 0; i < 9; ++i) qe[i] = std::strtod(argv[1 + i], nullptr);
    const auto Dn = getDn<N>();
    sri::ref::Matrix D_NN(n, n);
    for (int c = 0; c < 4; ++c) for (int i = 0; i < M; ++i) for (int j = 0; j < M; ++j) D_NN(c * M + i, c * M + j) = Dn(i, j);
    sri::ref::Matrix A_NN = D_NN;
    updateA<N, 3>(qe, A_NN, D_NN);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) std::printf("%.17g\\n", A_NN(i, j));
    sri::ref::Matrix bad = D_NN; bad(0, 0) += 1.0;
    try { updateA<N, 3>(qe, A_NN, bad); return 3; } catch (const std::invalid_argument&) {}
    return 0;
}
""")
    exe = tmp_path / "update_a"
    subprocess.run([gxx, "-std=c++17", "-O1", f"-I{ROOT / 'include'}", str(src), "-o", str(exe), f"-L{PKG}", "-lsri_cuda",
                    f"-Wl,-rpath,{PKG}"], check=True)
    out = subprocess.run([str(exe)] + [repr(float(v)) for v in qe], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout[-300:] + out.stderr
    A = np.array([float(v) for v in out.stdout.split()]).reshape(60, 60)
    gold = g["A_NN"][2]
    # entries updateA does not touch come from getDn<N>() (bit-identical to the reference's); the node-diagonal ones carry
    # K = Phi qe, which the device evaluates with fused multiply-adds: one ulp of |K| <= 4 is allowed there
    node_diag = np.kron(np.ones((4, 4)), np.eye(15)).astype(bool)
    assert np.array_equal(A[~node_diag], gold[~node_diag])
    assert np.abs(A - gold).max() <= 1e-15


# ---- error reporting ---------------------------------------------------------------------------------------------------

def test_singular_rod_is_reported_on_host_buffer_calls(h16, oracle16):
    """Host buffers + info array: every output lands, the call returns SRI_ERR_SINGULAR and info[] names the rods; the
    other rods are untouched by their neighbour's failure.  Without an info array the call stays SRI_OK."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SriError
    B = 300
    K, F, Mt, fb = oracle16.generate_rods(0x5EED, 9000, B)
    ref = oracle16.integrate_all(K, F, Mt, fbar=fb)
    Kbad = K.copy()
    Kbad[17, 0, 3] = np.nan
    info = np.full(B, -7, dtype=np.int32)
    Q = np.empty((B, 4, 15)); r = np.empty((B, 3, 15)); n = np.empty((B, 3, 15)); m = np.empty((B, 3, 15))
    with pytest.raises(SriError) as exc:
        h16.integrate_all(Kbad, F, Mt, fbar=fb, Q=Q, r=r, n=n, m=m, info=info)
    assert exc.value.code == -5 and "rod 17" in str(exc.value)
    good = np.ones(B, dtype=bool); good[17] = False
    assert info[17] != 0 and (info[good] == 0).all()
    for name, got in (("Q", Q), ("r", r), ("n", n), ("m", m)):
        assert rel_err(got[good], ref[name][good]) <= TOL, name
    out = h16.integrate_all(Kbad, F, Mt, fbar=fb)  # no info array: nothing to report through
    assert rel_err(out["Q"][good], ref["Q"][good]) <= TOL
    info[:] = -7
    h16.integrate_all(K, F, Mt, fbar=fb, Q=Q, r=r, n=n, m=m, info=info)  # all regular: SRI_OK
    assert (info == 0).all()


def test_handles_sharing_a_gpu_do_not_lower_each_others_shared_memory_limit(sri_lib, make_oracle, torch_mod):
    """cudaFuncAttributeMaxDynamicSharedMemorySize is per (kernel, device): handles with different N / optional inputs that
    alternate on one GPU must keep working (the limit was cached per handle).  N = 16 / N = 5 alternate on the
    TMA stage kernels, N = 64 / N = 40 on the Jacobian and the TMA stage kernels of the 64-row instantiation."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
    t = lambda a: torch_mod.from_numpy(np.ascontiguousarray(a)).cuda()
    for Na, Nb in ((16, 5), (64, 40)):
        oa, ob = make_oracle(Na), make_oracle(Nb)
        with SpectralRodIntegrator(Na, 0) as ha, SpectralRodIntegrator(Nb, 0) as hb:
            data = {}
            for h, o in ((ha, oa), (hb, ob)):
                B = 64
                K, F, Mt, fb = o.generate_rods(0x5EED, 10, B)
                rng = np.random.default_rng(o.N)
                lbar = rng.normal(size=(B, 3, o.N)); Gam = np.tile(np.array([1.0, 0.05, -0.02])[None, :, None], (B, 1, o.N))
                data[o.N] = (K, F, Mt, fb, lbar, Gam, o.integrate_all(K, F, Mt, fbar=fb, lbar=lbar, Gamma=Gam), o.integrate_all(K, F, Mt, fbar=fb))
            for rnd in range(3):
                for h, o in ((ha, oa), (hb, ob), (ha, oa)):
                    K, F, Mt, fb, lbar, Gam, ref_full, ref_plain = data[o.N]
                    Q = h.integrate_quaternions(t(K))
                    if rnd % 2 == 0:  # more optional inputs => larger staging layout
                        n = h.integrate_stress(t(F), fbar=t(fb))
                        m = h.integrate_couple(Q, n, t(Mt), Gamma=t(Gam), lbar=t(lbar))
                        r = h.integrate_position(Q, Gamma=t(Gam))
                        ref = ref_full
                    else:
                        n = h.integrate_stress(t(F), fbar=t(fb))
                        m = h.integrate_couple(Q, n, t(Mt))
                        r = h.integrate_position(Q)
                        ref = ref_plain
                    J = h.shape_jacobian(Q, n, m, t(Mt), 3, (1.0, 1.0, 0.77))
                    h.synchronize()
                    assert rel_err(m.cpu().numpy(), ref["m"]) <= TOL, (o.N, rnd)
                    assert rel_err(r.cpu().numpy(), ref["r"]) <= TOL, (o.N, rnd)
                    assert np.isfinite(J.cpu().numpy()).all()


def test_caller_device_is_restored(h16, torch_mod):
    """API calls make the handle's device current only for their own duration."""
    if torch_mod.cuda.device_count() < 2:
        # one GPU: the current device cannot differ from the handle's; still check that nothing changes it
        before = torch_mod.cuda.current_device()
        h16.synchronize()
        assert torch_mod.cuda.current_device() == before
        return
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
    torch_mod.cuda.set_device(1)
    try:
        with SpectralRodIntegrator(16, 0) as h:
            assert torch_mod.cuda.current_device() == 1
            K = torch_mod.zeros((4, 3, 16), dtype=torch_mod.float64, device="cuda:0")
            h.integrate_quaternions(K)
            h.synchronize()
            assert torch_mod.cuda.current_device() == 1
    finally:
        torch_mod.cuda.set_device(0)


def test_per_call_event_timers_and_nvtx_ranges(sri_lib, oracle16, torch_mod):
    """CUDA-event timers around each stage.  sri_set_timing brackets every call on the handle; the
    separate-stage entry points are thereby timed stage by stage, nested entry points (sri_integrate_quaternions ->
    sri_integrate_all) report the outer name once.  (The NVTX ranges are pushed unconditionally; they need a profiler to
    be seen, here we only check that pushing them is harmless.)"""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator, SriError
    B = 20000
    K, F, Mt, fb = oracle16.generate_rods(0x5EED, 0, B)
    t = lambda a: torch_mod.from_numpy(a).cuda()
    with SpectralRodIntegrator(16, 0) as h:
        with pytest.raises(SriError):
            h.last_timing()  # timing is off by default
        h.set_timing(True)
        with pytest.raises(SriError):
            h.last_timing()  # nothing timed yet
        dK, dF, dM, dfb = t(K), t(F), t(Mt), t(fb)
        seen = {}
        Q = h.integrate_quaternions(dK); seen["sri_integrate_quaternions"] = h.last_timing()
        r = h.integrate_position(Q); seen["sri_integrate_position"] = h.last_timing()
        n = h.integrate_stress(dF, fbar=dfb); seen["sri_integrate_stress"] = h.last_timing()
        m = h.integrate_couple(Q, n, dM); seen["sri_integrate_couple"] = h.last_timing()
        h.integrate_all(dK, dF, dM, fbar=dfb); seen["sri_integrate_all"] = h.last_timing()
        for name, (ms, got) in seen.items():
            assert got == name and 0.0 < ms < 50.0, (name, ms, got)
        # the quaternion stage is the expensive one (FP64-bound), the others stream
        assert seen["sri_integrate_quaternions"][0] > seen["sri_integrate_position"][0]
        h.set_timing(False)
        h.integrate_position(Q)
        with pytest.raises(SriError):
            h.last_timing()
        ref = oracle16.integrate_all(K[:100], F[:100], Mt[:100], fbar=fb[:100])
        assert rel_err(m[:100].cpu().numpy(), ref["m"]) <= TOL and rel_err(r[:100].cpu().numpy(), ref["r"]) <= TOL


# ---- several devices in one process -------------------------------------------------------------------------------------

def test_multi_device_host_path_is_bit_identical_to_one_handle(h16, oracle16, torch_mod):
    """sri_create_multi + sri_integrate_all_sharded (one host thread per shard).  On a one-GPU box the shards are two or
    three handles on device 0, which exercises the same code; results must equal the single-handle run bit for bit."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import MultiDeviceIntegrator, shard_range, SriError
    ndev = torch_mod.cuda.device_count()
    devices = list(range(ndev)) if ndev > 1 else [0, 0, 0]
    B = 70001  # ragged shards, more than one pipeline chunk per shard
    K, F, Mt, fb = oracle16.generate_rods(0x5EED, 31, B)
    one = h16.integrate_all(K, F, Mt, fbar=fb)
    with MultiDeviceIntegrator(16, devices=devices) as mh:
        assert mh.device_count == len(devices)
        info = np.full(B, -1, dtype=np.int32)
        got = mh.integrate_all(K, F, Mt, fbar=fb, info=info)
        assert (info == 0).all()
        for s in "Qrnm":
            assert np.array_equal(got[s], one[s]), s
        ref = oracle16.integrate_all(K[:500], F[:500], Mt[:500], fbar=fb[:500])
        for s in "Qrnm":
            assert rel_err(got[s][:500], ref[s]) <= TOL, s
        # a failure in one shard surfaces as the call's status, naming the device
        Kbad = K.copy(); Kbad[B - 5, 1, 1] = np.inf
        with pytest.raises(SriError) as exc:
            mh.integrate_all(Kbad, F, Mt, fbar=fb, info=info)
        assert exc.value.code == -5 and "device" in str(exc.value)
        lo, hi = shard_range(B, len(devices) - 1, len(devices))
        assert lo <= B - 5 < hi and info[B - 5] != 0 and (np.delete(info, B - 5) == 0).all()
        # sharded Newton == single-handle Newton (norms added in shard order: same history to round-off)
        Bn, ne, H = 5003, 3, (1.0, 1.0, 0.77)
        Fn = np.zeros((Bn, 3)); Fn[:, 2] = -np.linspace(0.1, 2.0, Bn); Mn = np.zeros((Bn, 3))
        q_m, rep_m = mh.newton_static_shape(Fn, Mn, ne, H)
        q_1, rep_1 = h16.newton_static_shape(Fn, Mn, ne, H, fd_step=0.0)
        assert rep_m["converged"] and rep_m["iterations"] == rep_1["iterations"]
        assert np.allclose(rep_m["rms_history"], rep_1["rms_history"], rtol=1e-9, atol=1e-16)
        assert np.array_equal(q_m, q_1)


def test_cpp_multi_gpu_example(sri_lib, torch_mod):
    exe = ROOT / "examples" / "multi_gpu_main_gpu"
    if not exe.exists():
        import __graft_entry__ as g
        g._build_cpp_example(PKG / "libsri_cuda.so")
    res = subprocess.run([str(exe), str(max(1, torch_mod.cuda.device_count())), "60000"], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "bit-identical" in res.stdout and "converged 1" in res.stdout


# ---- Newton driver: device-side reduction, lagged test, NCCL hook ---------------------------------------------------------

@pytest.mark.parametrize("fd_step", [0.0, 1e-6])
def test_lagged_device_side_convergence_test_takes_the_same_steps(h16, torch_mod, fd_step):
    """reduce == NULL: norms reduced and tested on the device, the host one iteration behind.  Iterates, iteration count and
    history must equal the synchronous loop (host callback = identity) exactly."""
    B, ne, H = 2311, 3, (1.0, 1.0, 0.77)
    rng = np.random.default_rng(5)
    F = np.zeros((B, 3)); F[:, 2] = -rng.uniform(0.1, 2.0, size=B)
    Mt = rng.uniform(-0.1, 0.1, size=(B, 3))
    tF, tM = (torch_mod.from_numpy(a).cuda() for a in (F, Mt))
    q_lag, rep_lag = h16.newton_static_shape(tF, tM, ne, H, fd_step=fd_step)
    q_syn, rep_syn = h16.newton_static_shape(tF, tM, ne, H, fd_step=fd_step, allreduce=lambda norms: None)
    assert rep_lag["converged"] and rep_syn["converged"]
    assert rep_lag["iterations"] == rep_syn["iterations"] and rep_lag["integrations"] == rep_syn["integrations"]
    assert rep_lag["rms_history"] == rep_syn["rms_history"]
    assert torch_mod.equal(q_lag, q_syn)
    assert rep_lag["singular_solves"] == 0
    # max_iter reached without convergence: same behaviour in both modes
    q_a, rep_a = h16.newton_static_shape(tF, tM, ne, H, fd_step=fd_step, max_iter=2)
    q_b, rep_b = h16.newton_static_shape(tF, tM, ne, H, fd_step=fd_step, max_iter=2, allreduce=lambda norms: None)
    assert not rep_a["converged"] and rep_a["iterations"] == 2 and rep_a["rms_history"] == rep_b["rms_history"]
    assert torch_mod.equal(q_a, q_b)


@pytest.mark.parametrize("fd_step", [0.0, 1e-6])
def test_newton_iteration_as_cuda_graph_replays_the_eager_loop(sri_lib, torch_mod, fd_step, monkeypatch):
    """From its second iteration on the lagged loop launches one CUDA graph per iteration (captured once per workspace, H,
    tolerance).  Same iterates, history and kernel count as eager launches (SRI_NEWTON_GRAPH=0); a changed H or tolerance
    re-captures instead of replaying stale arguments."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator, kernel_launch_count
    B, ne = 1500, 3
    rng = np.random.default_rng(77)
    F = np.zeros((B, 3)); F[:, 2] = -rng.uniform(0.1, 2.0, size=B)
    Mt = rng.uniform(-0.1, 0.1, size=(B, 3))
    tF, tM = (torch_mod.from_numpy(a).cuda() for a in (F, Mt))
    cases = [((1.0, 1.0, 0.77), 1e-10), ((1.0, 1.0, 0.77), 1e-10), ((2.0, 1.5, 0.9), 1e-10), ((2.0, 1.5, 0.9), 1e-6)]

    def run_all():
        out = []
        with SpectralRodIntegrator(16, 0) as h:
            for H, tol in cases:
                n0 = kernel_launch_count()
                q, rep = h.newton_static_shape(tF, tM, ne, H, fd_step=fd_step, tol=tol)
                out.append((q.cpu().numpy(), rep, kernel_launch_count() - n0))
        return out

    graph = run_all()
    monkeypatch.setenv("SRI_NEWTON_GRAPH", "0")
    eager = run_all()
    for (qg, rg, ng), (qe_, re_, ne_) in zip(graph, eager):
        assert rg["converged"] and rg["iterations"] >= 3
        assert rg["rms_history"] == re_["rms_history"] and rg["iterations"] == re_["iterations"]
        assert np.array_equal(qg, qe_)
        assert ng == ne_
    assert graph[3][1]["iterations"] < graph[2][1]["iterations"]   # the looser tolerance was honoured, not the captured one


def test_singular_newton_systems_are_counted_and_skipped(h16, torch_mod):
    """H = 0 and no load: residual and Jacobian vanish identically, every per-rod system is singular; the update must be
    skipped (qe stays finite) and the report must say so (NaN used to be written silently)."""
    B, ne = 64, 2
    F = torch_mod.zeros((B, 3), dtype=torch_mod.float64, device="cuda")
    qe0 = torch_mod.full((B, 3 * ne), 0.3, dtype=torch_mod.float64, device="cuda")
    K0 = torch_mod.ones((B, 3, 16), dtype=torch_mod.float64, device="cuda")
    # residual H (K - K0) with H = 0 is zero, so it would converge at once; use a tolerance of 0 to force iterations
    qe, rep = h16.newton_static_shape(F, F.clone(), ne, (0.0, 0.0, 0.0), qe=qe0.clone(), K0=K0, fd_step=0.0, tol=0.0, max_iter=2)
    assert torch_mod.isfinite(qe).all() and torch_mod.equal(qe, qe0)
    assert rep["singular_solves"] == 2 * B and rep["iterations"] == 2


def test_nccl_hook_single_rank(h16, torch_mod):
    """sri_nccl_unique_id / sri_nccl_init / sri_nccl_allreduce_norms with a one-rank communicator (libnccl is opened at run
    time): the reduction is the identity and the Newton solve takes the same steps with and without the communicator.
    (Two and more ranks: bench.py's cfg5 leg under torchrun.)"""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator
    B, ne, H = 500, 3, (1.0, 1.0, 0.77)
    F = torch_mod.zeros((B, 3), dtype=torch_mod.float64, device="cuda"); F[:, 2] = -torch_mod.linspace(0.1, 2.0, B, dtype=torch_mod.float64)
    Mt = torch_mod.zeros_like(F)
    q_ref, rep_ref = h16.newton_static_shape(F, Mt, ne, H, fd_step=0.0)
    with SpectralRodIntegrator(16, 0) as h:
        h.nccl_init(1, 0, h.nccl_unique_id())
        norms = torch_mod.tensor([3.5, 0.25], dtype=torch_mod.float64, device="cuda")
        h.nccl_allreduce_norms(norms)
        h.synchronize()
        assert norms.tolist() == [3.5, 0.25]
        q, rep = h.newton_static_shape(F, Mt, ne, H, fd_step=0.0)
        assert rep["rms_history"] == rep_ref["rms_history"] and torch_mod.equal(q, q_ref)
        h.nccl_finalize()
        q2, rep2 = h.newton_static_shape(F, Mt, ne, H, fd_step=0.0)
        assert torch_mod.equal(q2, q_ref)


# ---- rod length (rod_modeling.pdf eq. 2.17) -------------------------------------------------------------------

@pytest.mark.parametrize("N,ell", [(16, 0.37), (16, 2.5), (32, 1.8)])
def test_rod_length_scaling(sri_lib, make_oracle, torch_mod, N, ell):
    """A rod of length ell: every ODE of the normalised problem is multiplied by ell (eq. 2.17), i.e. the library is called
    with (ell K, ell Gamma, ell fbar, ell lbar) -- scale_for_length().  Checked against an independent numpy collocation on
    the PHYSICAL interval [0, ell] (nodes ComputeChebyshevPoints<N, L = ell>, differentiation matrix Dn / ell)."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator, scale_for_length
    o = make_oracle(N)
    M = N - 1
    rng = np.random.default_rng(N)
    B = 6
    x = o.chebyshev_points(1.0)
    K = rng.uniform(-1.5, 1.5, size=(B, 3, 1)) + rng.uniform(-1, 1, size=(B, 3, 1)) * (2 * x - 1)[None, None, :]
    Gam = np.tile(np.array([1.0, 0.03, -0.02])[None, :, None], (B, 1, N)) + 0.01 * rng.normal(size=(B, 3, N))
    fb = rng.normal(size=(B, 3, 1)) * np.ones((1, 1, N)); lb = 0.3 * rng.normal(size=(B, 3, 1)) * np.ones((1, 1, N))
    F = rng.uniform(-1, 1, size=(B, 3)); Mt = rng.uniform(-1, 1, size=(B, 3))
    t = lambda a: torch_mod.from_numpy(np.ascontiguousarray(a)).cuda()
    sK, sG, sf, sl = scale_for_length(ell, K, Gam, fb, lb)
    with SpectralRodIntegrator(N, 0) as h:
        got = {k: v.cpu().numpy() for k, v in h.integrate_all(t(sK), t(F), t(Mt), Gamma=t(sG), fbar=t(sf), lbar=t(sl)).items()}
        # the C-ABI form (sri_scale_for_length) scales in place: device tensors with one length per rod, host arrays with a scalar
        dK, dG, df, dl = t(K), t(Gam), t(fb), t(lb)
        h.scale_for_length_(torch_mod.full((B,), ell, dtype=torch_mod.float64, device="cuda"), dK, dG, df, dl)
        h.synchronize()
        for got_t, want in ((dK, sK), (dG, sG), (df, sf), (dl, sl)):
            assert np.array_equal(got_t.cpu().numpy(), want)
        hK, hf = K.copy(), fb.copy()
        h.scale_for_length_(ell, K=hK, fbar=hf)
        assert np.array_equal(hK, sK) and np.array_equal(hf, sf)
    D = o.dn() / ell  # d/ds on [0, ell]
    for b in range(B):
        A = np.kron(np.eye(4), D[:M, :M])
        for i in range(M):
            k0, k1, k2 = K[b, :, i]
            Ak = np.array([[0, -k0, -k1, -k2], [k0, 0, k2, -k1], [k1, -k2, 0, k0], [k2, k1, -k0, 0]])
            for rr in range(4):
                for cc in range(4):
                    A[rr * M + i, cc * M + i] -= 0.5 * Ak[rr, cc]
        Q = np.linalg.solve(A, -np.kron(np.array([1.0, 0, 0, 0]), D[:M, M]))
        assert np.abs(Q - got["Q"][b].reshape(-1)).max() <= 1e-11
        q = np.concatenate([Q.reshape(4, M), np.array([[1.0], [0], [0], [0]])], axis=1)  # all N nodes
        w, xq, y, z = q
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (xq * y - w * z), 2 * (xq * z + w * y)],
                      [2 * (xq * y + w * z), 1 - 2 * (xq * xq + z * z), 2 * (y * z - w * xq)],
                      [2 * (xq * z - w * y), 2 * (y * z + w * xq), 1 - 2 * (xq * xq + y * y)]])  # [3][3][N]
        rp = np.einsum("ijn,jn->in", R, Gam[b])  # dr/ds at all nodes
        r = np.linalg.solve(D[:M, :M], rp[:, :M].T)
        assert np.abs(r.T - got["r"][b]).max() <= 1e-11
        n = np.linalg.solve(D[1:, 1:], -fb[b][:, 1:].T - np.outer(D[1:, 0], F[b]))
        assert np.abs(n.T - got["n"][b]).max() <= 1e-11
        rhs = -(np.cross(rp[:, 1:].T, n) + lb[b][:, 1:].T) - np.outer(D[1:, 0], Mt[b])
        m = np.linalg.solve(D[1:, 1:], rhs)
        assert np.abs(m.T - got["m"][b]).max() <= 1e-10 * max(1.0, np.abs(m).max())


# ---- two ranks, two GPUs: the NCCL hook end to end (skipped on one-GPU boxes) -----------------------------------------------

def _nccl_rank(rank, world, id_path, out_path, B, ne):
    import time

    import torch

    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator, shard_range
    torch.cuda.set_device(rank)
    h = SpectralRodIntegrator(16, rank)
    if rank == 0:
        uid = h.nccl_unique_id()
        Path(id_path + ".tmp").write_bytes(uid)
        Path(id_path + ".tmp").rename(id_path)
    else:
        while not Path(id_path).exists():
            time.sleep(0.01)
        uid = Path(id_path).read_bytes()
    h.nccl_init(world, rank, uid)
    lo, hi = shard_range(B, rank, world)
    F = torch.zeros((hi - lo, 3), dtype=torch.float64, device=f"cuda:{rank}")
    F[:, 2] = -torch.linspace(0.1, 2.0, B, dtype=torch.float64)[lo:hi].to(F.device)
    Mt = torch.zeros_like(F)
    qe, rep = h.newton_static_shape(F, Mt, ne, (1.0, 1.0, 0.77), fd_step=0.0, total_dof=3 * ne * B)
    norms = torch.tensor([float(rank + 1), float(10 - rank)], dtype=torch.float64, device=f"cuda:{rank}")
    h.nccl_allreduce_norms(norms)
    h.synchronize()
    np.savez(out_path + f".{rank}.npz", qe=qe.cpu().numpy(), hist=np.array(rep["rms_history"]), it=rep["iterations"],
             conv=rep["converged"], norms=norms.cpu().numpy())
    h.nccl_finalize()
    h.close()


def test_nccl_norm_reduction_two_ranks(sri_lib, h16, torch_mod, tmp_path):
    """sri_nccl_init over two ranks (one process per GPU, the unique id exchanged through a file): the sharded Newton solve
    takes the steps of the single-rank solve, both ranks see the same residual history, and sri_nccl_allreduce_norms is
    (sum, max) over the ranks."""
    if torch_mod.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    B, ne = 4001, 3
    id_path, out_path = str(tmp_path / "nccl_id"), str(tmp_path / "out")
    mp.spawn(_nccl_rank, args=(2, id_path, out_path, B, ne), nprocs=2, join=True)
    parts = [np.load(out_path + f".{r}.npz") for r in range(2)]
    F = torch_mod.zeros((B, 3), dtype=torch_mod.float64, device="cuda"); F[:, 2] = -torch_mod.linspace(0.1, 2.0, B, dtype=torch_mod.float64)
    q1, rep1 = h16.newton_static_shape(F, torch_mod.zeros_like(F), ne, (1.0, 1.0, 0.77), fd_step=0.0)
    assert all(bool(p["conv"]) for p in parts) and all(int(p["it"]) == rep1["iterations"] for p in parts)
    assert np.array_equal(parts[0]["hist"], parts[1]["hist"])            # folded in rank order: identical bits on every rank
    assert np.allclose(parts[0]["hist"], rep1["rms_history"], rtol=1e-9, atol=1e-18)
    assert np.array_equal(np.concatenate([p["qe"] for p in parts]), q1.cpu().numpy())
    for p in parts:
        assert p["norms"].tolist() == [3.0, 10.0]
