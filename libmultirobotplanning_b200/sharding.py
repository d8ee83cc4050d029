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
