"""The lane-exact numpy model of the DMMA Gauss-Jordan (tools/dmma_gj_emulator.py) against the oracle: pins the index
algebra of csrc/sri_fused16_dmma.cuh (fragment layouts, shuffle sources, sign tables, pivot-row trick) on the CPU."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))
import dmma_gj_emulator as emu  # noqa: E402


@pytest.mark.parametrize("N", [16, 11, 3])
def test_emulated_dmma_elimination_matches_oracle(make_oracle, N):
    o = make_oracle(N)
    M = N - 1
    S = o.operator(3)
    g = -S @ o.operator(2)
    Stx = emu.tables(S, g, M)
    K, _, _, _ = o.generate_rods(0x5EED, 42, 5)
    rng = np.random.default_rng(N)
    q0 = rng.normal(size=(5, 4)); q0 /= np.linalg.norm(q0, axis=1, keepdims=True)
    ref = o.integrate_all(K, q0=q0, want=("Q",))["Q"]
    for b in range(5):
        Q, flagged = emu.solve_rod(Stx, K[b], q0[b], M, growth=4.0)
        assert not flagged
        assert np.abs(Q.T - ref[b]).max() <= 1e-12 * np.abs(ref[b]).max()


def test_growth_check_flags_a_tiny_bound(make_oracle):
    o = make_oracle(16)
    S = o.operator(3)
    Stx = emu.tables(S, -S @ o.operator(2), 15)
    K, _, _, _ = o.generate_rods(0x5EED, 7, 1)
    _, flagged = emu.solve_rod(Stx, K[0], np.array([1.0, 0, 0, 0]), 15, growth=1e-6)
    assert flagged
