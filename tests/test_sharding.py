"""Host-side multi-GPU logic on CPU: the contiguous batch partition and the optional final gather, run with
world_size = 2 over `gloo` (no GPU; the per-rank "solve" is the CPU oracle standing in for the device)."""
import os
import socket

import numpy as np
import pytest

from solvempc_b200.sharding import gather_results, shard_arrays, shard_bounds


def test_shard_bounds_partition_the_batch():
    for batch in (0, 1, 7, 4096, 65537):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(batch, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _worker(rank, world, port, batch, out_dir):
    import torch.distributed as dist
    import oracle
    from problems import c2_batch
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cfg = oracle.load_config(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config", "MPC_API.json"))
        m = oracle.mpc_build(**cfg)
        X, U, ref = c2_batch(batch, seed=11)                      # every rank builds the same batch deterministically ...
        Xs, Us, rs = shard_arrays((X, U, ref), world, rank)       # ... and owns a contiguous shard of it
        f, ub = oracle.mpc_batch_vectors(m, Xs, Us, rs)
        out = oracle.solve_batch(m["H"], m["Gbar"], m["lb"], m["W0"], f, ub, nthreads=1, eps_abs=1e-5, eps_rel=1e-5)
        u0 = gather_results(Us + out["x"][:, 0], batch)           # optional final gather of the controls
        st = gather_results(out["status"], batch)
        np.save(os.path.join(out_dir, f"u0_{rank}.npy"), u0)
        np.save(os.path.join(out_dir, f"st_{rank}.npy"), st)
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_reproduce_the_single_process_batch(tmp_path):
    import torch.multiprocessing as mp
    import oracle
    from problems import c2_batch
    batch, world = 37, 2                                          # ragged: 19 + 18
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(world, port, batch, str(tmp_path)), nprocs=world, join=True)
    cfg = oracle.load_config(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "config", "MPC_API.json"))
    m = oracle.mpc_build(**cfg)
    X, U, ref = c2_batch(batch, seed=11)
    f, ub = oracle.mpc_batch_vectors(m, X, U, ref)
    one = oracle.solve_batch(m["H"], m["Gbar"], m["lb"], m["W0"], f, ub, nthreads=1, eps_abs=1e-5, eps_rel=1e-5)
    for rank in range(world):
        u0 = np.load(tmp_path / f"u0_{rank}.npy")
        st = np.load(tmp_path / f"st_{rank}.npy")
        assert np.array_equal(u0, U + one["x"][:, 0])             # bit-identical: sharding changes nothing but the owner
        assert np.array_equal(st, one["status"])
