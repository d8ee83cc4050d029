"""Distance fields of narrow maps (dimx <= 30, more than one tile high: one bitmap word per row before the
fix of bitmapRowWords) against the oracle.  usage: python tests/narrow_check.py  (a checker script, not a pytest file: it uses the oracle, hence it lives under tests/)"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
t0 = time.time()
from libmultirobotplanning_b200 import capi  # noqa: E402
from oracle import orc  # noqa: E402

capi.init(0)
rng = np.random.default_rng(5)
bad = 0
for dimx, dimy in ((17, 45), (30, 33), (1, 40), (8, 100), (3, 64), (29, 1000), (20, 40), (31, 33), (64, 5)):
    for dens in (0.0, 0.2, 0.4):
        blocked = rng.random((dimy, dimx)) < dens
        ys, xs = np.nonzero(blocked)
        obst = np.stack([xs, ys], 1).astype(np.int32)
        cells = rng.choice(dimx * dimy, min(dimx * dimy, 9), replace=False)
        goals = np.stack([cells % dimx, cells // dimx], 1).astype(np.int32)
        want = orc.bfs_fields(dimx, dimy, obst, goals)
        for env in ({}, {"MRP_BFS_TILES": "1"}):
            os.environ.update(env)
            got = capi.bfs_fields(dimx, dimy, obst, goals)
            for k in env:
                os.environ.pop(k)
            n = int((got != want).sum())
            bad += n
            print("%4dx%-4d d=%.1f %-6s mismatching cells: %d" % (dimx, dimy, dens, "tiles" if env else "queue", n), flush=True)
print("TOTAL mismatches", bad, "in %.1f s" % (time.time() - t0), flush=True)
