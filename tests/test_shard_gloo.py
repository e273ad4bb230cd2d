"""CPU test of the multi-GPU host logic: problems shard contiguously over ranks with no collective in
the data path; the only exchange is the final gather of U.  Runs the real torch.distributed plumbing that bench.py
uses, with the gloo backend and world_size 2 (the per-rank "solver" is a stand-in: the GPU kernels need a GPU)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import ROOT


def test_shard_range_partitions_exactly():
    from bench_problems import shard_range
    for total in (0, 1, 7, 4096, 2 ** 20, 1000003):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(total, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    from bench_problems import shard_range
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    B, M = 64, 6
    rng = np.random.default_rng(11)           # one seeded stream shared by all ranks, as bench.py does
    X_all = rng.standard_normal((B * world, 5)).astype(np.float32)
    lo, hi = shard_range(B * world, world, rank)
    X = X_all[lo:hi]
    U_local = torch.from_numpy(np.tanh(X @ np.arange(30, dtype=np.float32).reshape(5, M)))  # stand-in for the solve
    U_all = torch.empty((B * world, M), dtype=torch.float32)
    dist.all_gather_into_tensor(U_all, U_local)   # the one collective of the path
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)      # max-over-ranks timing reduction
    dist.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, "U_all.npy"), U_all.numpy())
        np.save(os.path.join(out_dir, "tmax.npy"), t.numpy())
    dist.destroy_process_group()


def test_world_size_2_gather(tmp_path):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    U_all = np.load(tmp_path / "U_all.npy")
    rng = np.random.default_rng(11)
    X_all = rng.standard_normal((128, 5)).astype(np.float32)
    want = np.tanh(X_all @ np.arange(30, dtype=np.float32).reshape(5, 6))
    np.testing.assert_allclose(U_all, want, rtol=1e-6)
    assert np.load(tmp_path / "tmax.npy")[0] == 2.0
