"""How the hot path shards over the GPUs of one box.

Distance fields shard by goal, conflict checks by table / agent-pair block and
replans by job: units are independent, so every rank processes its slice with
no data-path collective.  The only collective is the optional all-gather that
makes every goal's field resident on every GPU (north_star)."""


def shard_range(n_units, rank, world):
    """Contiguous, balanced slice [begin, end) of n_units for `rank`."""
    base, rem = divmod(n_units, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_goals(goal_xy, rank, world):
    b, e = shard_range(len(goal_xy), rank, world)
    return goal_xy[b:e]


def allgather_fields(local_fields, n_total, rank, world, dist):
    """All-gathers per-rank field blocks [n_local, cells] (torch tensors on the
    backend's device) into [n_total, cells] in goal order.  Uneven shards are
    padded to the largest one, as torch.distributed requires equal sizes."""
    import torch
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    n_max = max(e - b for b, e in sizes)
    cells = local_fields.shape[1]
    padded = torch.zeros((n_max, cells), dtype=local_fields.dtype, device=local_fields.device)
    padded[:local_fields.shape[0]] = local_fields
    gathered = torch.empty((world * n_max, cells), dtype=local_fields.dtype,
                           device=local_fields.device)
    dist.all_gather_into_tensor(gathered, padded)
    parts = [gathered[r * n_max:r * n_max + (e - b)] for r, (b, e) in enumerate(sizes)]
    return torch.cat(parts, 0)


# ---- conflict detection sharded by time slab -------------------------------
# getFirstConflict / focalHeuristic walk t = 0 .. max_t-1 and test every agent
# pair at every step (example/cbs.cpp:335-386, ecbs.cpp:315-350); steps are
# independent, so rank r takes the steps [t0, t1) of its slab on a copy of the
# path table, and the results meet in two all-reduces: MIN over the packed
# first-conflict keys (t, type, i, j — the reference's iteration order) and SUM
# over the counts.  The path table itself is broadcast beforehand.

NO_CONFLICT = -1                      # the kernels' "no conflict" key, as int64
KEY_T_SHIFT = 41                      # key = t << 41 | type << 40 | i << 20 | j


def conflict_time_range(length, mode):
    """Loop bound of the reference: max(len) - 1 for cbs / ecbs (mode 0,
    cbs.cpp:338-341), max(len) for cbs_ta (mode 1, cbs_ta.cpp:372-375)."""
    m = int(length.max()) if len(length) else 0
    return max(m - 1, 0) if mode == 0 else m


def slab_table(table, length, t0, t1):
    """Path table of the steps [t0, t1): columns pos_i(t0) .. pos_i(t1) with the
    reference's clamp to the last state (getState, cbs.cpp:420-429) applied, and
    uniform lengths t1 - t0 + 1, so that a mode-0 sweep over it visits exactly
    the steps of the slab.  Works on torch tensors and numpy arrays."""
    n = t1 - t0 + 1
    if hasattr(table, "gather"):  # torch
        import torch
        idx = torch.arange(t0, t1 + 1, device=table.device).unsqueeze(0).expand(table.shape[0], n)
        idx = torch.minimum(idx, (length.to(torch.int64) - 1).clamp(min=0).unsqueeze(1))
        sub = table.gather(1, idx).contiguous()
        return sub, torch.full((table.shape[0],), n, dtype=length.dtype, device=length.device)
    import numpy as np
    idx = np.minimum(np.arange(t0, t1 + 1)[None, :], np.maximum(length.astype(np.int64) - 1, 0)[:, None])
    sub = np.ascontiguousarray(np.take_along_axis(table, idx, 1))
    return sub, np.full(table.shape[0], n, length.dtype)


def shift_key(key, t0):
    """First-conflict key of a slab (local step) -> key of the whole table."""
    return key if key == NO_CONFLICT else key + (t0 << KEY_T_SHIFT)


def reduce_conflicts(key, count, dist, device=None):
    """MIN of the first-conflict keys (NO_CONFLICT = none) and SUM of the counts
    over the ranks."""
    import torch
    big = torch.iinfo(torch.int64).max
    k = torch.tensor([big if key == NO_CONFLICT else key], dtype=torch.int64, device=device)
    c = torch.tensor([count], dtype=torch.int64, device=device)
    dist.all_reduce(k, op=dist.ReduceOp.MIN)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    k = int(k.item())
    return (NO_CONFLICT if k == big else k), int(c.item())
