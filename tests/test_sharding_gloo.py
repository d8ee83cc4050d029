"""CPU test of the N>1 path: world_size 2 over gloo.  Each rank computes the
distance fields of its goal shard (with the oracle standing in for the GPU
kernel — this test is about the sharding and the all-gather, not the kernel)
and the gathered result must equal the single-process answer."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_goals, out_path):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from libmultirobotplanning_b200 import instances, sharding
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    inst = instances.load_set(os.path.join(ROOT, "tests/golden/bench_32x32.npz"))[900]
    goals = inst.goals[:n_goals]
    mine = sharding.shard_goals(goals, rank, world)
    local = torch.from_numpy(orc.bfs_fields(inst.dimx, inst.dimy, inst.obstacles, mine))
    full = sharding.allgather_fields(local, n_goals, rank, world, dist)
    if rank == 0:
        np.save(out_path, full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_partition():
    from libmultirobotplanning_b200.sharding import shard_range
    for n in (0, 1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            cuts = [shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[r][1] == cuts[r + 1][0] for r in range(world - 1))
            sizes = [e - b for b, e in cuts]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.parametrize("n_goals", [7, 10])
def test_two_rank_allgather_matches_single(tmp_path, n_goals, orc):
    import torch.multiprocessing as mp
    from libmultirobotplanning_b200 import instances
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "full.npy")
    mp.spawn(_worker, args=(2, port, n_goals, out), nprocs=2, join=True)
    inst = instances.load_set(os.path.join(ROOT, "tests/golden/bench_32x32.npz"))[900]
    want = orc.bfs_fields(inst.dimx, inst.dimy, inst.obstacles, inst.goals[:n_goals])
    assert np.array_equal(np.load(out), want)


def _key(c):
    return -1 if c is None else (c[0] << 41) | (c[3] << 40) | (c[1] << 20) | c[2]


def _random_table(seed, n=24, tmax=30, cells=40):
    rng = np.random.default_rng(seed)
    length = rng.integers(1, tmax + 1, n).astype(np.int32)
    tpad = int(length.max())
    cell = np.zeros((n, tpad), np.int32)
    for i in range(n):
        p = int(rng.integers(0, cells))
        for t in range(tpad):
            if t < length[i]:
                p = int(np.clip(p + rng.integers(-1, 2), 0, cells - 1))
            cell[i, t] = p           # padded by repeating the last cell
    return cell, length


def _conflict_worker(rank, world, port, seeds, out_path):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from libmultirobotplanning_b200 import sharding
    from oracle import orc
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    res = []
    for seed in seeds:
        for mode in (0, 1):
            cell, length = _random_table(seed)
            t_end = sharding.conflict_time_range(length, mode)
            t0, t1 = sharding.shard_range(t_end, rank, world)
            key, count = sharding.NO_CONFLICT, 0
            if t1 > t0:
                sub, sublen = sharding.slab_table(cell, length, t0, t1)
                # a slab is always swept with the cbs/ecbs bound: its last column is
                # only there for the edge test of the last step
                key = sharding.shift_key(_key(orc.first_conflict(sub, sublen, 64, 0)), t0)
                count = orc.count_conflicts(sub, sublen, 0)
            res.append(sharding.reduce_conflicts(key, count, dist))
    if rank == 0:
        np.save(out_path, np.array(res, np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_conflicts_sharded_by_time_slab(tmp_path, world, orc):
    """Each rank sweeps the steps of its slab (oracle standing in for the
    kernel); MIN over the keys and SUM over the counts must equal the answers on
    the whole table, for both loop bounds of the reference."""
    import torch.multiprocessing as mp
    seeds = (1, 2, 3, 4)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "c.npy")
    mp.spawn(_conflict_worker, args=(world, port, seeds, out), nprocs=world, join=True)
    got = np.load(out).reshape(len(seeds), 2, 2)
    for k, seed in enumerate(seeds):
        cell, length = _random_table(seed)
        for mode in (0, 1):
            assert got[k, mode, 0] == _key(orc.first_conflict(cell, length, 64, mode)), (seed, mode)
            assert got[k, mode, 1] == orc.count_conflicts(cell, length, mode), (seed, mode)
