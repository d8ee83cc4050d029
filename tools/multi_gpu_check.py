"""Multi-GPU parity + timing of the product's sharded operations (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py [--goals G] [--time]
Checks, on every rank: mrp_bfs_fields_allgather_dev (fields by goal slice + packed NCCL all-gather + device
expansion) == mrp_bfs_fields_dev of all goals on this GPU, on the C5 map and (int32 fall-back) on a maze;
mrp_conflicts_sharded_dev (agent-pair blocks + all-reduce) == mrp_conflicts_dev.  torch.distributed only
carries the 128-byte communicator id."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import libmultirobotplanning_b200 as pkg

capi = pkg.capi
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
G = int(sys.argv[sys.argv.index("--goals") + 1]) if "--goals" in sys.argv else 600
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
capi.init(local)
if world > 1:
    dist.init_process_group("gloo")
    idt = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8)
    if rank == 0:
        idt = torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8).clone()
    dist.broadcast(idt, 0)
    capi.comm_init_rank(bytes(idt.numpy().tobytes()), world, rank)
info = capi.comm_info()
assert info["n_ranks"] == world and info["rank"] == rank, info
stream = torch.cuda.current_stream()


def fields_both_ways(mp, goal_cells, cells):
    n = len(goal_cells)
    d_goals = torch.from_numpy(goal_cells).to(dev)
    ref = torch.empty((n, cells), dtype=torch.int32, device=dev)
    ws = torch.empty(max(mp.workspace_bytes(n), 256), dtype=torch.uint8, device=dev)
    mp.bfs_fields_dev(d_goals.data_ptr(), n, ref.data_ptr(), ws.data_ptr(), stream.cuda_stream)
    out = torch.full((n, cells), -7, dtype=torch.int32, device=dev)
    ws2 = torch.empty(max(mp.allgather_workspace_bytes(n), 256), dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(3):
        out.fill_(-7)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        mp.bfs_fields_allgather_dev(d_goals.data_ptr(), n, out.data_ptr(), ws2.data_ptr(), stream.cuda_stream)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    same = bool(torch.equal(out, ref))
    return same, min(ts), capi.comm_last_gather(), ref


# (1) C5 map
inst = pkg.instances.synthetic_c5(n_agents=max(G, 64))
mp = capi.Map(1024, 1024, inst.obstacles)
gc = (inst.goals[:G, 0] + 1024 * inst.goals[:G, 1]).astype(np.int32)
same, ms, st, ref = fields_both_ways(mp, gc, 1 << 20)
print("rank %d/%d: C5 %d goals gathered == single-GPU: %s; %.2f ms; %s" % (rank, world, G, same, ms, st), flush=True)
assert same
assert world == 1 or st["bytes_per_cell"] == 1
if "--fields-only" in sys.argv:
    if world > 1:
        dist.barrier()
        capi.comm_destroy()
        dist.destroy_process_group()
    sys.exit(0)
# (2) a maze: detours beyond a byte -> int32 fall-back
dimx, dimy = 256, 128
obst = []
for y in range(1, dimy, 2):
    gap = dimx - 1 if (y // 2) % 2 == 0 else 0
    obst += [[x, y] for x in range(dimx) if x != gap]
mz = capi.Map(dimx, dimy, np.asarray(obst, np.int32))
rng = np.random.default_rng(3)
gcm = rng.choice(dimx * dimy, 37, replace=False).astype(np.int32)
same, ms, st, _ = fields_both_ways(mz, gcm, dimx * dimy)
print("rank %d: maze 37 goals gathered == single-GPU: %s (wire format %d B/cell)" % (rank, same, st["bytes_per_cell"]), flush=True)
assert same
assert world == 1 or st["bytes_per_cell"] == 4
# (3) conflicts by agent-pair block
N, T = 700, 150
cell = rng.integers(0, 3000, (N, T)).astype(np.int32)
ln = rng.integers(1, T + 1, N).astype(np.int32)
d_cell, d_len = torch.from_numpy(cell).to(dev), torch.from_numpy(ln).to(dev)
lib = capi.lib()
for mode in (0, 1):
    for first, count in ((1, 1), (1, 0), (0, 1)):
        a = torch.zeros(4, dtype=torch.int64, device=dev)
        b = torch.zeros(4, dtype=torch.int64, device=dev)
        capi.check(lib.mrp_conflicts_dev(d_cell.data_ptr(), d_len.data_ptr(), N, T, mode, first, count, a.data_ptr(), stream.cuda_stream))
        capi.check(lib.mrp_conflicts_sharded_dev(d_cell.data_ptr(), d_len.data_ptr(), N, T, mode, first, count, b.data_ptr(), stream.cuda_stream))
        torch.cuda.synchronize()
        if first:
            assert int(a[0]) == int(b[0]), (mode, first, count, a, b)
        if count:
            assert int(a[1]) == int(b[1]), (mode, first, count, a, b)
print("rank %d: sharded conflict sweep == single-GPU sweep" % rank, flush=True)
if world > 1:
    dist.barrier()
    capi.comm_destroy()
    dist.destroy_process_group()
print("rank %d OK" % rank, flush=True)
