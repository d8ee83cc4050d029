import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libmultirobotplanning_b200 as pkg
from oracle import orc
pkg.capi.init(0)
s32 = pkg.instances.load_set(os.path.join(ROOT, "tests/golden/bench_32x32.npz"))
s8 = pkg.instances.load_set(os.path.join(ROOT, "tests/golden/bench_8x8.npz"))
for n_agents, n_inst in ((50, 16), (100, 16)):
    insts = [i for i in s32 if i.n_agents == n_agents][:n_inst]
    t = time.time()
    res = pkg.solver.solve_batch(pkg.solver.ECBS, insts, w=1.3, max_hl=3000, max_seconds=120)
    dt = time.time() - t
    ok = [r for r in res if r["status"] == 0]
    print("GPU ECBS w=1.3 %d agents: %d/%d solved in %.2fs (%.2f inst/s) hl=%s ll=%d cost/LB=%s" % (
        n_agents, len(ok), len(insts), dt, len(ok) / dt, [r["hl_expanded"] for r in res][:8],
        sum(r["ll_expanded"] for r in res),
        ["%.3f" % (r["cost"] / r["lower_bound"]) for r in ok][:6]))
    t = time.time()
    n_cpu = 4
    cres = [orc.ecbs(i.dimx, i.dimy, i.obstacles, i.starts, i.goals, 1.3, (3000, 0, 60.0)) for i in insts[:n_cpu]]
    dt = time.time() - t
    print("  CPU oracle: %d/%d solved in %.2fs (%.2f inst/s) hl=%s costs gpu/cpu %s" % (
        sum(r["status"] == 0 for r in cres), n_cpu, dt, sum(r["status"] == 0 for r in cres) / dt,
        [r["hl_expanded"] for r in cres], [(a["cost"], b["cost"]) for a, b in zip(res, cres)]))
t = time.time()
res = pkg.solver.solve_batch(pkg.solver.CBS, s8, max_hl=2000, max_seconds=120)
dt = time.time() - t
print("GPU CBS 8x8 full set: %d/2000 solved in %.2fs, hl total %d, ll total %d" % (
    sum(r["status"] == 0 for r in res), dt, sum(r["hl_expanded"] for r in res), sum(r["ll_expanded"] for r in res)))
t = time.time()
cres = [orc.cbs(i.dimx, i.dimy, i.obstacles, i.starts, i.goals, (2000, 0, 5.0)) for i in s8[::10]]
dt = time.time() - t
print("CPU oracle CBS 8x8 every 10th: %d/200 solved in %.2fs, hl %d" % (sum(r["status"] == 0 for r in cres), dt, sum(r["hl_expanded"] for r in cres)))
