"""Per-warp busy cycles of the sweep BFS kernel (CTA 0).
Build first: tools/build_variant.sh timing -DMRP_SWEEP_TIMING bfs_sweep.cu
run: MRP_B200_LIB=build/timing/libmrp_b200.so python tools/sweep_timing.py"""
import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import libmultirobotplanning_b200 as pkg
capi = pkg.capi; capi.init(0)
inst = pkg.instances.synthetic_c5(n_agents=256)
G = 148
gc = (inst.goals[:G, 0] + 1024 * inst.goals[:G, 1]).astype(np.int32)
mp = capi.Map(1024, 1024, inst.obstacles)
d_goals = torch.from_numpy(gc).cuda()
d_out = torch.empty((G, 1 << 20), dtype=torch.int32, device="cuda")
ws = torch.empty(mp.workspace_bytes(G), dtype=torch.uint8, device="cuda")
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mp.bfs_fields_dev(d_goals.data_ptr(), G, d_out.data_ptr(), ws.data_ptr(), 0)
    e1.record()
    torch.cuda.synchronize()
    print("148 goals: %.3f ms" % e0.elapsed_time(e1))
b = np.zeros(40, np.uint64)
capi.lib().mrp_debug_sweep_busy(b.ctypes.data_as(ctypes.c_void_p))
goals, steps, total = int(b[33]), int(b[32]), int(b[34])
print("CTA 0: %d goals, %d steps, %.0f cycles per goal, %.0f cycles per step" % (goals, steps, total / goals, total / steps))
for w in range(32):
    print("warp %2d: busy %6.0f cycles per step" % (w, b[w] / steps))
