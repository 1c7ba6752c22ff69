"""Host-side multi-rank logic on CPU: rod-index sharding and the residual-norm all-reduce (gloo, world_size 2)."""
import os
import socket

import numpy as np
import pytest

from experimental_gpu_programming_for_a_spectral_numerical_integration_b200.sharding import allreduce_residual, shard_range


def test_shard_ranges_tile_the_batch():
    for total in (0, 1, 7, 1000, 10 ** 6 + 3):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_range(total, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == total
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def test_c_abi_shard_range_equals_the_python_one():
    """sri_shard_range (what a C/C++ host and sri_integrate_all_sharded use) cuts exactly where sharding.shard_range does;
    pure host arithmetic, no GPU needed."""
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import SriError
    from experimental_gpu_programming_for_a_spectral_numerical_integration_b200 import shard_range as c_shard_range
    for total in (0, 1, 7, 1000, 10 ** 6 + 3, 2 ** 62 + 12345):
        for world in (1, 2, 3, 8, 64):
            for rank in range(world):
                assert c_shard_range(total, rank, world) == shard_range(total, rank, world)
    with pytest.raises(SriError):
        c_shard_range(10, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, total, out_dir):
    import torch
    import torch.distributed as dist

    from oracle.oracle import Oracle

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        o = Oracle(16)
        lo, hi = shard_range(total, rank, world)
        K, F, Mt, fb = o.generate_rods(0x5EED, lo, hi - lo)          # every rank regenerates only its own rods
        out = o.integrate_all(K, F, Mt, fbar=fb)
        rho = o.shape_residual(K, np.array([1.0, 1.0, 0.77]), out["Q"], out["m"], Mt)
        red = torch.tensor([float((rho ** 2).sum()), float(np.abs(rho).max())], dtype=torch.float64)
        allreduce_residual(red)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), Q=out["Q"], m=out["m"], red=red.numpy(), lo=lo, hi=hi)
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(tmp_path):
    import torch.multiprocessing as mp

    from oracle.oracle import Oracle, build_oracle

    build_oracle()
    total, world = 301, 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, total, str(tmp_path)), nprocs=world, join=True)
    o = Oracle(16)
    K, F, Mt, fb = o.generate_rods(0x5EED, 0, total)
    ref = o.integrate_all(K, F, Mt, fbar=fb)
    rho = o.shape_residual(K, np.array([1.0, 1.0, 0.77]), ref["Q"], ref["m"], Mt)
    parts = [np.load(tmp_path / f"rank{r}.npz") for r in range(world)]
    assert np.array_equal(np.concatenate([p["Q"] for p in parts]), ref["Q"])     # bit-identical to the unsharded run
    assert np.array_equal(np.concatenate([p["m"] for p in parts]), ref["m"])
    for p in parts:
        assert abs(p["red"][0] - (rho ** 2).sum()) <= 1e-12 * (rho ** 2).sum()
        assert p["red"][1] == np.abs(rho).max()
    assert np.array_equal(parts[0]["red"], parts[1]["red"])  # one all-gather folded in rank order: identical bits everywhere
