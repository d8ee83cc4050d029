"""Tiny invocation of every kernel, for compute-sanitizer (one tool per call)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libmultirobotplanning_b200 as pkg
capi = pkg.capi; capi.init(0)
rng = np.random.default_rng(0)
blocked = rng.random((70, 90)) < 0.2
ys, xs = np.nonzero(blocked); obst = np.stack([xs, ys], 1)
f = capi.bfs_fields(90, 70, obst, [[3, 3], [80, 60]])            # bfs_tiles_kernel
f2 = capi.bfs_fields(20, 20, obst[(obst[:, 0] < 20) & (obst[:, 1] < 20)], [[1, 1]])  # bfs_small
N, T = 300, 6
a = rng.integers(0, 200, (N, 1)); cell = np.clip(a + np.cumsum(rng.integers(-1, 2, (N, T)), 1), 0, 199).astype(np.int32)
ln = np.full(N, T, np.int32)
print("hashed", capi.first_conflict(cell, ln, 32, 0), capi.count_conflicts(cell, ln, 0))
print("pairs", capi.count_conflicts(cell[:40], ln[:40], 1))
m = capi.Map(20, 20, obst[(obst[:, 0] < 20) & (obst[:, 1] < 20)])
fld = capi.bfs_fields(20, 20, obst[(obst[:, 0] < 20) & (obst[:, 1] < 20)], [[18, 18]])
free = np.flatnonzero(fld[0] != capi.INF)
r = capi.lowlevel_batch([m], fld, [{"start": int(free[0]), "goal": 18 + 20 * 18, "field": 0, "vc": [(3, int(free[5]))]}],
                        w=1.3, tables=cell[:8, :].reshape(1, 8, T) % 400, table_len=ln[:8].reshape(1, 8), max_expanded=500)
print("lowlevel", r[0]["status"], r[0]["cost"])
