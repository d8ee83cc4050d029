import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libmultirobotplanning_b200 as pkg
from oracle import orc
capi = pkg.capi; capi.init(0)
rng = np.random.default_rng(7)
bad = 0
for trial in range(60):
  for N, cells in ((4000, 30000), (4096, 50000), (3900, 8000)):
    a = rng.integers(0, cells, N).astype(np.int32)
    b = np.clip(a + rng.integers(-1, 2, N), 0, cells - 1).astype(np.int32)
    cell = np.stack([a, b], 1).astype(np.int32); ln = np.full(N, 2, np.int32)
    o = orc.count_conflicts(cell, ln, 0)
    of = orc.first_conflict(cell, ln, 1000, 0)
    hs = [capi.count_conflicts(cell, ln, 0) for _ in range(4)]
    fs = [capi.first_conflict(cell, ln, 1000, 0) for _ in range(2)]
    if any(h != o for h in hs) or any(f != of for f in fs):
        bad += 1
        print(trial, N, "oracle", o, "hashed", hs, of, fs)
print("bad trials:", bad, "of 180")
