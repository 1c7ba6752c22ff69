"""Builds the reference's OWN sources verbatim (TEST INFRASTRUCTURE ONLY).

/root/reference/main.cpp + include/*.h are compiled where they lie against oracle/eigen_shim (a minimal stand-in for
the Eigen/Boost API subset they use; neither library is in the image).  Outputs, both under oracle/_ref/ (git-ignored,
shipped to the GPU box with the snapshot):
  reference_main             the reference's executable, verbatim
  libreference_harness.so    oracle/reference_harness.cpp, which #includes the reference's main.cpp (main() renamed away) and
                             calls its own integrateQuaternions() / integratePosition() / updateA over a batch of rods
Nothing from /root/reference is copied into the repository.

The reference's own build system (CMake + find_package(Eigen3)) is not run: it cannot succeed here.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
from pathlib import Path

import numpy as np

ORACLE_DIR = Path(__file__).resolve().parent
REF_SRC = Path("/root/reference")
REF_BIN = ORACLE_DIR / "_ref" / "reference_main"
REF_HARNESS = ORACLE_DIR / "_ref" / "libreference_harness.so"


def build_reference(force: bool = False):
    """Returns the binary path, or None when /root/reference is not present (e.g. on the GPU box)."""
    main_cpp = REF_SRC / "main.cpp"
    if not main_cpp.exists():
        return REF_BIN if REF_BIN.exists() else None
    if REF_BIN.exists() and not force and REF_BIN.stat().st_mtime >= max(
            p.stat().st_mtime for p in (ORACLE_DIR / "eigen_shim").rglob("*") if p.is_file()):
        return REF_BIN
    REF_BIN.parent.mkdir(exist_ok=True)
    gxx = shutil.which("g++") or "g++"
    cmd = [gxx, "-std=c++20", "-O2", "-I", str(ORACLE_DIR / "eigen_shim"), "-I", str(REF_SRC / "include"),
           str(main_cpp), "-o", str(REF_BIN)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the reference against the shim failed:\n" + res.stdout + res.stderr)
    return REF_BIN


def build_reference_harness(force: bool = False):
    """oracle/reference_harness.cpp + /root/reference/main.cpp -> oracle/_ref/libreference_harness.so.
    Returns the library path, or None when neither the sources nor a prebuilt library are available."""
    main_cpp = REF_SRC / "main.cpp"
    src = ORACLE_DIR / "reference_harness.cpp"
    if not main_cpp.exists():
        return REF_HARNESS if REF_HARNESS.exists() else None
    deps = [src, main_cpp] + [p for p in (ORACLE_DIR / "eigen_shim").rglob("*") if p.is_file()]
    if REF_HARNESS.exists() and not force and REF_HARNESS.stat().st_mtime >= max(p.stat().st_mtime for p in deps):
        return REF_HARNESS
    REF_HARNESS.parent.mkdir(exist_ok=True)
    gxx = shutil.which("g++") or "g++"
    # -ffp-contract=off: plain IEEE arithmetic, as the portable oracle build (no FMA contraction of the reference's loops)
    cmd = [gxx, "-std=c++20", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", str(ORACLE_DIR / "eigen_shim"),
           "-I", str(REF_SRC / "include"), "-I", str(REF_SRC), str(src), "-o", str(REF_HARNESS)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("building the reference harness against the shim failed:\n" + res.stdout + res.stderr)
    return REF_HARNESS


class ReferenceHarness:
    """The reference's own functions (N = 16, ne = na = 3 as compiled into main.cpp) over a batch of rods."""

    N = 16

    def __init__(self):
        import ctypes
        lib = build_reference_harness()
        if lib is None or not Path(lib).exists():
            raise FileNotFoundError("reference harness not available (needs /root/reference or a prebuilt oracle/_ref)")
        self.lib = ctypes.CDLL(str(lib))
        self.lib.sri_reference_integrate_rods.restype = ctypes.c_int
        self.lib.sri_reference_integrate_rods.argtypes = [ctypes.c_long] + [ctypes.c_void_p] * 4
        self.lib.sri_reference_update_A.argtypes = [ctypes.c_void_p] * 2

    def integrate(self, qe: np.ndarray):
        """qe [n][9] -> dict(K [n][3][16], Q [n][4][15], r [n][3][15]) straight from main.cpp's functions."""
        qe = np.ascontiguousarray(qe, dtype=np.float64).reshape(-1, 9)
        n, N, M = qe.shape[0], self.N, self.N - 1
        K = np.empty((n, 3, N)); Q = np.empty((n, 4, M)); r = np.empty((n, 3, M))
        got = self.lib.sri_reference_integrate_rods(n, qe.ctypes.data, K.ctypes.data, Q.ctypes.data, r.ctypes.data)
        assert got == N
        return {"K": K, "Q": Q, "r": r}

    def update_A(self, qe: np.ndarray) -> np.ndarray:
        """A_NN (60 x 60) as updateA (main.cpp:55-88) leaves it for one qe (9,)."""
        qe = np.ascontiguousarray(qe, dtype=np.float64).reshape(9)
        n = 4 * (self.N - 1)
        buf = np.empty(n * n)
        self.lib.sri_reference_update_A(qe.ctypes.data, buf.ctypes.data)
        return buf.reshape(n, n).T.copy()


def run_reference(precision: int | None = 17):
    """Runs the reference's main() and parses its two dumps.  precision=None keeps its native 6 digits."""
    exe = build_reference()
    if exe is None or not Path(exe).exists():
        raise FileNotFoundError("reference binary not available")
    env = dict(os.environ)
    if precision is None:
        env.pop("SRI_SHIM_PRECISION", None)
    else:
        env["SRI_SHIM_PRECISION"] = str(precision)
    out = subprocess.run([str(exe)], capture_output=True, text=True, env=env, check=True).stdout
    m = re.match(r"Q_stack : \n(.*)\nr_stack : \n(.*)\n", out, re.S)
    if not m:
        raise RuntimeError("unexpected reference output:\n" + out[:400])
    Q = np.array([float(v) for v in m.group(1).split()])
    r = np.array([[float(v) for v in line.split()] for line in m.group(2).strip().split("\n")])
    return {"stdout": out, "Q_stack": Q, "r_stack": r}


if __name__ == "__main__":
    print(build_reference(force=True))
    print(build_reference_harness(force=True))
