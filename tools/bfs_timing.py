"""Per-warp phase timing of the tiled BFS kernel (needs -DMRP_BFS_TIMING)."""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import libmultirobotplanning_b200 as pkg
capi = pkg.capi; capi.init(0)
inst = pkg.instances.synthetic_c5(n_agents=64)
G = 148
gc = (inst.goals[:1, 0] + 1024 * inst.goals[:1, 1]).astype(np.int32).repeat(G)
mp = capi.Map(1024, 1024, inst.obstacles)
d_goals = torch.from_numpy(gc).cuda()
d_out = torch.empty((G, 1 << 20), dtype=torch.int32, device="cuda")
ws = torch.empty(mp.workspace_bytes(G), dtype=torch.uint8, device="cuda")
for _ in range(2):
    mp.bfs_fields_dev(d_goals.data_ptr(), G, d_out.data_ptr(), ws.data_ptr(), 0)
torch.cuda.synchronize()
t = np.zeros((32, 12), np.uint64)
capi.lib().mrp_debug_bfs_timing(t.ctypes.data_as(ctypes.c_void_p))
lv = 1498
print("levels", lv)
print("warp  loopTop  phaseA  wait1  B:post-stores->bar  wait2 | entries | B:claims  B:alloc  B:stores")
for w in range(32):
    if t[w].sum() == 0: continue
    print(w, " ".join("%7.0f" % (float(x) / lv) for x in t[w][:5]), "   %.1f  |" % (float(t[w][5]) / lv), " ".join("%7.0f" % (float(x) / lv) for x in t[w][6:9]))
