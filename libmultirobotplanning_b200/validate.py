"""Validity of a multi-agent solution under the reference's own conflict
semantics, vectorised (numpy) so that whole batches of 100-200-agent solutions
can be checked: every path starts on its start at g = 0 and ends on its goal,
moves one free cell at a time, and the clamped path table has no vertex conflict
and no edge swap inside the loop bound of the reference's getFirstConflict
(example/cbs.cpp:335-386: t < max(len) - 1 for cbs / ecbs, mode 0;
example/cbs_ta.cpp:369-420: t < max(len), mode 1)."""
import numpy as np


def path_table(paths, dimx):
    """cells x + dimx * y of the paths, clamped to the last state like getState
    (example/cbs.cpp:420-429) -> int64 [N][T]."""
    T = max(len(p) for p in paths)
    tab = np.empty((len(paths), T), np.int64)
    for a, p in enumerate(paths):
        p = np.asarray(p)
        c = p[:, 0].astype(np.int64) + dimx * p[:, 1].astype(np.int64)
        tab[a, :len(c)] = c
        tab[a, len(c):] = c[-1]
    return tab


def validate_paths(inst, paths, mode=0, goals=None):
    """Returns None if the solution is valid, else a string saying what is wrong.
    goals: [N][2] final cells to check (defaults to inst.goals when present)."""
    free = np.ones((inst.dimy, inst.dimx), bool)
    obst = np.asarray(inst.obstacles).reshape(-1, 2)
    if len(obst):
        ok = (obst[:, 0] >= 0) & (obst[:, 0] < inst.dimx) & (obst[:, 1] >= 0) & (obst[:, 1] < inst.dimy)
        free[obst[ok, 1], obst[ok, 0]] = False
    if goals is None and getattr(inst, "goals", None) is not None:
        goals = inst.goals
    starts = np.asarray(inst.starts).reshape(-1, 2)
    if len(paths) != len(starts):
        return "%d paths for %d agents" % (len(paths), len(starts))
    for a, p in enumerate(paths):
        p = np.asarray(p)
        if len(p) == 0:
            return "agent %d: empty path" % a
        if tuple(p[0][:2]) != tuple(starts[a]) or p[0][2] != 0:
            return "agent %d: does not start on its start at g = 0" % a
        if goals is not None and tuple(p[-1][:2]) != tuple(np.asarray(goals)[a]):
            return "agent %d: does not end on its goal" % a
        x, y = p[:, 0], p[:, 1]
        if (x < 0).any() or (y < 0).any() or (x >= inst.dimx).any() or (y >= inst.dimy).any():
            return "agent %d: leaves the map" % a
        if not free[y, x].all():
            return "agent %d: steps on an obstacle" % a
        if len(p) > 1 and (np.abs(np.diff(x)) + np.abs(np.diff(y)) > 1).any():
            return "agent %d: moves more than one cell" % a
    assert inst.dimx * inst.dimy <= 1 << 21
    tab = path_table(paths, inst.dimx)
    T = tab.shape[1] - (1 if mode == 0 else 0)  # timesteps the reference tests
    if T <= 0:
        return None
    # positions at T are needed for the swap test of step T - 1
    ext = np.concatenate([tab, tab[:, -1:]], 1) if T == tab.shape[1] else tab
    cols = np.sort(ext[:, :T], axis=0)
    hit = np.nonzero((cols[1:] == cols[:-1]).any(axis=0))[0]
    if len(hit):
        return "vertex conflict at t = %d" % int(hit[0])
    a, b = ext[:, :T], ext[:, 1:T + 1]
    t = np.broadcast_to(np.arange(T, dtype=np.int64), a.shape)
    moving = a != b
    fwd = (t[moving] << 42) | (a[moving] << 21) | b[moving]
    rev = (t[moving] << 42) | (b[moving] << 21) | a[moving]
    both = np.intersect1d(fwd, rev)
    if len(both):
        return "edge swap at t = %d" % int(both[0] >> 42)
    return None
