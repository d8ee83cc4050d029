"""Driver for ncu: the conflict sweep of config C5 (N agents on their goal fields).
usage: python tools/prof_conflicts.py [n_agents]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import libmultirobotplanning_b200 as pkg  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
capi = pkg.capi
capi.init(0)
inst = pkg.instances.synthetic_c5(n_agents=N)
DIM = 1024
gc = (inst.goals[:N, 0] + DIM * inst.goals[:N, 1]).astype(np.int32)
mp = capi.Map(DIM, DIM, inst.obstacles)
d_goals = torch.from_numpy(gc).cuda()
d_out = torch.empty((N, DIM * DIM), dtype=torch.int32, device="cuda")
ws = torch.empty(max(mp.workspace_bytes(N), 256), dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream()
mp.bfs_fields_dev(d_goals.data_ptr(), N, d_out.data_ptr(), ws.data_ptr(), s.cuda_stream)
starts = (inst.starts[:, 0] + DIM * inst.starts[:, 1]).astype(np.int64)
table, length = bench.descend_paths(torch, d_out, inst, starts, N, 4096)
d_res = torch.zeros(4, dtype=torch.int64, device="cuda")
lib = capi.lib()
print("sum of (len - 1) = %d moving agent-steps of %d x %d" % (int((length.clamp(min=1) - 1).sum().item()), N,
                                                               int(length.max().item()) - 1))
for r in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    capi.check(lib.mrp_conflicts_dev(table.data_ptr(), length.data_ptr(), N, table.shape[1], 0, 1, 1,
                                     d_res.data_ptr(), s.cuda_stream))
    e1.record()
    torch.cuda.synchronize()
    print("sweep %d: %.3f ms, count %d, max_t %d" % (r, e0.elapsed_time(e1), int(d_res[1].item()),
                                                    int(length.max().item()) - 1))
first = []
for r in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    capi.check(lib.mrp_conflicts_dev(table.data_ptr(), length.data_ptr(), N, table.shape[1], 0, 1, 0,
                                     d_res.data_ptr(), s.cuda_stream))
    e1.record()
    torch.cuda.synchronize()
    first.append(e0.elapsed_time(e1))
print("first conflict only: %s ms, key %d" % (" ".join("%.3f" % x for x in first), int(d_res[0].item())))
