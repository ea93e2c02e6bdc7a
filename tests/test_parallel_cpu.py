"""Host-side logic of the point-sharded multi-GPU path, on CPU with gloo (world_size 2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dbat_b200.parallel import exchange_unique_id, partition_points
from dbat_b200.synth import make_scene


def test_partition_covers_and_balances():
    rng = np.random.default_rng(0)
    nOP = 1000
    op = np.sort(rng.integers(0, nOP, 20000))
    for world in (1, 2, 3, 8):
        parts = partition_points(op, nOP, world)
        assert parts[0][0] == 0 and parts[-1][1] == nOP
        assert all(parts[r][1] == parts[r + 1][0] for r in range(world - 1))
        cnt = [np.count_nonzero((op >= lo) & (op < hi)) for lo, hi in parts]
        assert sum(cnt) == len(op)
        assert max(cnt) - min(cnt) <= 0.05 * len(op) / world + 60


def test_shards_partition_observations_and_unknowns():
    """Every observation and every OP unknown belongs to exactly one shard; camera unknowns are shared."""
    s, _ = make_scene(12, 300, rays=4, seed=2)
    parts = partition_points(s.IP.op, s.OP.val.shape[1], 3)
    des = s.bundle.deserial.OP
    seen_obs = np.zeros(len(s.IP.op), int)
    seen_cols = np.zeros(s.bundle.serial.n, int)
    for lo, hi in parts:
        seen_obs[(s.IP.op >= lo) & (s.IP.op < hi)] += 1
        seen_cols[des.src[(des.dest >= 3 * lo) & (des.dest < 3 * hi)]] += 1
    nC = s.bundle.serial.n - len(s.bundle.serial.OP.dest)
    assert np.all(seen_obs == 1)
    assert np.all(seen_cols[nC:] == 1) and np.all(seen_cols[:nC] == 0)


def _free_port():
    with socket.socket() as so:
        so.bind(('127.0.0.1', 0))
        return so.getsockname()[1]


def _worker(rank, world, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    uid = exchange_unique_id(rank, lambda: bytes(range(128)), dist)
    # the allreduce pattern of ShardedProblem.gather: owners contribute their columns, rank 0 the camera part
    n, nC = 20, 6
    own = np.arange(nC + rank, n, world)
    v = np.full(n, float(rank + 1))
    t = torch.zeros(n, dtype=torch.float64)
    t[torch.from_numpy(own)] = torch.from_numpy(v[own])
    if rank == 0:
        t[:nC] = torch.from_numpy(v[:nC])
    dist.all_reduce(t)
    out[rank] = (uid, t.numpy().copy())
    dist.destroy_process_group()


def test_unique_id_exchange_and_gather_gloo():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert out[0][0] == out[1][0] == bytes(range(128))
    expect = np.empty(20)
    expect[:6] = 1.0
    for r in range(world):
        expect[np.arange(6 + r, 20, world)] = r + 1
    np.testing.assert_array_equal(out[0][1], expect)
    np.testing.assert_array_equal(out[1][1], expect)
