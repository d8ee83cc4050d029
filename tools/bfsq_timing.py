"""Per-warp phase timing of the queue BFS kernel.
Build first with: make -C libmultirobotplanning_b200/csrc clean all EXTRA=-DMRP_BFS_TIMING"""
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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mp.bfs_fields_dev(d_goals.data_ptr(), G, d_out.data_ptr(), ws.data_ptr(), 0)
    e1.record()
    torch.cuda.synchronize()
    print("148 x same goal: %.3f ms" % e0.elapsed_time(e1))
lvl = np.zeros((4096, 2), np.uint32)
capi.lib().mrp_debug_bfsq_levels(lvl.ctypes.data_as(ctypes.c_void_p))
n = int((lvl[:, 0] > 0).sum()) + 1
print('levels', n)
cnt = lvl[2:n, 0].astype(np.float64)
dt = (lvl[3:n + 1, 1].astype(np.int64) - lvl[2:n, 1].astype(np.int64)) & 0xffffffff
A = np.stack([np.ones_like(cnt), cnt], 1)
coef = np.linalg.lstsq(A, dt.astype(np.float64), rcond=None)[0]
print("level cycles ~ %.0f + %.2f * count" % (coef[0], coef[1]))
for lo, hi in [(1, 64), (64, 128), (128, 256), (256, 384), (384, 512), (512, 768), (768, 1024), (1024, 1536), (1536, 4096)]:
    m = (cnt >= lo) & (cnt < hi)
    if m.any():
        print("count %4d-%4d: %4d levels, mean cycles %.0f" % (lo, hi, m.sum(), dt[m].mean()))

c0, c1, c2 = int(lvl[0, 0]), int(lvl[1, 0]), int(lvl[1, 1])
first, last = int(lvl[2, 1]), int(lvl[n - 1, 1])
m = 0xffffffff
print("cycles: goal start -> first level %d, level loop %d, loop end -> sweep start %d, sweep %d" % (
    (first - c0) & m, (last - first) & m, (c1 - last) & m, (c2 - c1) & m))

bl = np.zeros((1024, 4), np.uint64)
capi.lib().mrp_debug_bfsq_blocks(bl.ctypes.data_as(ctypes.c_void_p))
bl = bl[:G].astype(np.int64)
init = (bl[:, 2] - bl[:, 0]) / 1e3
loop = (bl[:, 3] - bl[:, 2]) / 1e3
sweep = (bl[:, 1] - bl[:, 3]) / 1e3
for name, v in (("init", init), ("level loop", loop), ("tail rings+exit", sweep), ("total", init + loop + sweep)):
    print("%-10s min %.1f  p25 %.1f  median %.1f  p75 %.1f  max %.1f us" % (name, v.min(), np.percentile(v, 25), np.median(v), np.percentile(v, 75), v.max()))
order = np.argsort(loop)
print("fastest loop blocks", order[:8], "slowest", order[-8:])
