import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_cuda() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


HAS_CUDA = _has_cuda()


def pytest_collection_modifyitems(config, items):
    if HAS_CUDA:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle16():
    from oracle import Oracle, build_oracle
    build_oracle()
    return Oracle(16)


@pytest.fixture(scope="session")
def make_oracle():
    from oracle import Oracle, build_oracle
    build_oracle()
    cache = {}

    def get(N):
        if N not in cache:
            cache[N] = Oracle(N)
        return cache[N]

    return get


@pytest.fixture(scope="session")
def sri_lib():
    """The CUDA library, built if needed (nvcc cross-compiles on CPU).  Never falls back to anything else."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200.build import build_library
    build_library()
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import _lib
    return _lib.load()


DEFAULT_QE = np.array([0, 0, 0, 1.2877691307032, -1.63807499160786, 0.437406679142598, 0, 0, 0], dtype=np.float64)


def rel_err(x, ref):
    """Per-rod relative error: max|x-ref| / max|ref| over each rod's stack; returns the worst rod."""
    x = np.asarray(x).reshape(x.shape[0], -1)
    ref = np.asarray(ref).reshape(ref.shape[0], -1)
    num = np.abs(x - ref).max(axis=1)
    den = np.maximum(np.abs(ref).max(axis=1), 1e-300)
    return float((num / den).max())
