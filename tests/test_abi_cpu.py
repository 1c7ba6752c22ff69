"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every symbol include/sri.h
declares; the host-side operator functions match the oracle bit for bit; nothing in the product touches oracle/."""
import ctypes
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import HAS_CUDA

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "experimental_gpu_programming_for_a_spectral_numerical_integration_b200"


def _declared_symbols():
    text = (ROOT / "include" / "sri.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sri_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_bound_and_exported(sri_lib):
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert sorted(_lib.SYMBOLS) == declared, "ctypes table and include/sri.h disagree"
    nm = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (sri_[A-Za-z0-9_]+)", nm))
    assert set(declared) <= exported


def test_library_contains_sm100a_code(sri_lib):
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_host_operator_functions_match_oracle_bitwise(sri_lib, make_oracle):
    import experimental_gpu_programming_for_a_spectral_numerical_integration_b200 as sri
    for N in (2, 3, 8, 16, 32, 64):
        o = make_oracle(N)
        assert np.array_equal(sri.ComputeChebyshevPoints(N), o.chebyshev_points())
        assert np.array_equal(sri.GetCoefficients_c(N), o.coefficients_c())
        assert np.array_equal(sri.getDn(N), o.dn())
    assert np.array_equal(sri.ComputeChebyshevPoints(16, 2.0), make_oracle(16).chebyshev_points(2.0))
    for X in (0.0, 0.25, 0.9):
        assert np.array_equal(sri.Phi(3, 3, X), make_oracle(16).phi(3, 3, X))
        assert np.array_equal(sri.Phi(2, 5, X, -1.0, 2.0), make_oracle(16).phi(2, 5, X, -1.0, 2.0))


def test_argument_errors_are_reported_not_thrown(sri_lib):
    assert sri_lib.sri_chebyshev_dn(1, None) == -1
    assert b"sri_chebyshev_dn" in sri_lib.sri_last_error_string()
    h = ctypes.c_void_p()
    assert sri_lib.sri_create(65, 0, ctypes.byref(h)) == -2 and not h
    assert sri_lib.sri_create(16, 0, None) == -1
    assert sri_lib.sri_destroy(None) == 0
    # every entry point that takes a handle rejects a null handle with a status, also the ones added for the static shape
    # problem (no compute is attempted)
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import _lib
    H = (ctypes.c_double * 3)(1.0, 1.0, 0.77)
    rep = _lib.NewtonReport()
    assert sri_lib.sri_newton_static_shape(None, 0, 3, H, None, None, None, None, 1e-10, 30, 1e-6, 0, _lib.ALLREDUCE_FN(), None,
                                           ctypes.byref(rep)) == -1
    assert sri_lib.sri_galerkin_residual(None, 0, 3, None, None, H, None, None, None, None, None, None) == -1
    assert sri_lib.sri_generalised_forces(None, 0, 3, None, None) == -1
    assert sri_lib.sri_integrate_wrench_local(None, 0, None, None, None, None, None, None, None, None, None, None) == -1
    assert b"handle" in sri_lib.sri_last_error_string().lower()


@pytest.mark.skipif(HAS_CUDA, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_a_gpu(sri_lib):
    """The product path must fail loudly when there is no CUDA device."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SpectralRodIntegrator, SriError
    with pytest.raises(SriError) as e:
        SpectralRodIntegrator(16, 0)
    assert e.value.code == -3


def test_product_never_touches_the_oracle():
    for path in list(PKG.rglob("*.py")) + list(PKG.rglob("*.cu")) + list(PKG.rglob("*.cuh")) + list(PKG.rglob("*.hpp")) + \
            [ROOT / "include" / "sri.h", ROOT / "include" / "sri_reference_api.hpp"]:
        if not path.exists():
            continue
        text = path.read_text()
        assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), path
        assert "sri_oracle" not in text and "libsri_oracle" not in text, path


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", tmp_path / "libsri_cuda.so")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load()


def test_skew_and_ad_helpers():
    """skew / ad of include/utilities.h:16-37 (spec source of the statics stages): Python mirror and C++ header."""
    import numpy as np
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import ad, skew
    rng = np.random.default_rng(3)
    v, w, g = rng.normal(size=3), rng.normal(size=3), rng.normal(size=3)
    assert np.allclose(skew(v) @ w, np.cross(v, w), atol=0, rtol=1e-15)
    assert np.array_equal(skew(v).T, -skew(v))
    A = ad(np.concatenate([v, g]))
    assert np.array_equal(A[:3, :3], skew(v)) and np.array_equal(A[3:, 3:], skew(v))
    assert np.array_equal(A[3:, :3], skew(g)) and not A[:3, 3:].any()


def test_cpp_header_skew_ad_compile_and_agree(tmp_path):
    import shutil, subprocess
    from pathlib import Path
    gxx = shutil.which("g++")
    if gxx is None:
        import pytest
        pytest.skip("no g++")
    root = Path(__file__).resolve().parent.parent
    src = tmp_path / "t.cpp"
    src.write_text("""#include "sri_reference_api.hpp"
#include <cstdio>
int main() {
    auto A = sri::ref::ad({1, 2, 3, 4, 5, 6});
    // [[k^,0],[g^,k^]]: A(0,1) = -k3, A(3,1) = -g3, A(4,3) = k3, A(0,4) = 0
    std::printf("%g %g %g %g\\n", A(0, 1), A(3, 1), A(4, 3), A(0, 4));
    return 0;
}
""")
    exe = tmp_path / "t"
    subprocess.run([gxx, "-std=c++17", f"-I{root / 'include'}", "-fsyntax-only", str(src)], check=True)
    # link-free check of the values: compile with the library only when it exists (sri.h symbols are not used here)
    res = subprocess.run([gxx, "-std=c++17", f"-I{root / 'include'}", str(src), "-o", str(exe),
                          f"-L{root / 'experimental_gpu_programming_for_a_spectral_numerical_integration_b200'}", "-lsri_cuda",
                          f"-Wl,-rpath,{root / 'experimental_gpu_programming_for_a_spectral_numerical_integration_b200'}"],
                         capture_output=True, text=True)
    if res.returncode == 0:
        out = subprocess.run([str(exe)], capture_output=True, text=True)
        if out.returncode == 0:
            assert out.stdout.split() == ["-3", "-6", "3", "0"]
