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
