"""Small driver for ncu: runs the tiled BFS kernel on the C5 map for a few goals.
usage: python tools/prof_bfs.py [n_goals] [reps]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import libmultirobotplanning_b200 as pkg  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else 296
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
capi = pkg.capi
capi.init(0)
inst = pkg.instances.synthetic_c5(n_agents=max(G, 64))
cells = 1024 * 1024
gc = (inst.goals[:G, 0] + 1024 * inst.goals[:G, 1]).astype(np.int32)
mp = capi.Map(1024, 1024, inst.obstacles)
d_goals = torch.from_numpy(gc).cuda()
d_out = torch.empty((G, cells), dtype=torch.int32, device="cuda")
ws = torch.empty(max(mp.workspace_bytes(G), 256), dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream()
for r in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mp.bfs_fields_dev(d_goals.data_ptr(), G, d_out.data_ptr(), ws.data_ptr(), s.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    print("rep %d: %d goals %.3f ms  (%.1f us/goal/SM)" % (r, G, e0.elapsed_time(e1),
          e0.elapsed_time(e1) * 1e3 * 148 / G))
