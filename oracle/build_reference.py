"""Builds the reference's OWN sources verbatim (TEST INFRASTRUCTURE ONLY).

/root/reference/main.cpp + include/*.h are compiled where they lie against oracle/eigen_shim (a minimal stand-in for
the Eigen/Boost API subset they use; neither library is in the image).  Output: oracle/_ref/reference_main (git-ignored,
shipped to the GPU box with the snapshot).  Nothing from /root/reference is copied into the repository.

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
