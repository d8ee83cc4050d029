"""Launch-by-launch profile of the ECBS batch of bench.py (config C3, rank 0 shard).
usage: MRP_LL_PROFILE=1 MRP_HOST_PROFILE=1 python tools/prof_ecbs.py [n_instances] 2> log"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import libmultirobotplanning_b200 as pkg  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
pkg.capi.init(0)
s32 = pkg.instances.load_set(os.path.join(ROOT, "tests", "golden", "bench_32x32.npz"))
insts, _ = bench.c3_shard(pkg, s32, 0)
insts = insts[:n]
pkg.solver.solve_batch(pkg.solver.ECBS, insts[:2], w=1.3, max_hl=50)
sys.stderr.write("==== timed batch ====\n")
runs = int(os.environ.get("PROF_RUNS", "1"))  # later runs find the lanes' device arenas grown
for _ in range(runs):
    t0 = time.perf_counter()
    res = pkg.solver.solve_batch(pkg.solver.ECBS, insts, w=1.3, max_hl=2000, max_seconds=120)
    dt = time.perf_counter() - t0
    if runs > 1:
        print("run: %.2f s" % dt, end="; ")
ok = sum(r["status"] == 0 for r in res)
print("%d/%d solved in %.2f s = %.1f instances/s" % (ok, len(insts), dt, ok / dt))
import collections
print("sum of costs of the solved instances", sum(r["cost"] for r in res if r["status"] == 0),
      " sum of hl_expanded", sum(r["hl_expanded"] for r in res))
print("status histogram", collections.Counter(r["status"] for r in res))
hl = sorted(r["hl_expanded"] for r in res)
print("hl_expanded: median %d p90 %d p99 %d max %d" % (hl[len(hl) // 2], hl[int(len(hl) * .9)], hl[int(len(hl) * .99)], hl[-1]))

if os.environ.get("PROF_CBS"):
    s8 = pkg.instances.load_set(os.path.join(ROOT, "tests", "golden", "bench_8x8.npz"))
    pkg.solver.solve_batch(pkg.solver.CBS, s8[:4], max_hl=50)
    t0 = time.perf_counter()
    res = pkg.solver.solve_batch(pkg.solver.CBS, s8, max_hl=500, max_seconds=120)
    dt = time.perf_counter() - t0
    print("CBS 8x8: %d/%d solved in %.2f s, %.3g HL expansions/s, cost sum %d" % (
        sum(r["status"] == 0 for r in res), len(s8), dt, sum(r["hl_expanded"] for r in res) / dt,
        sum(r["cost"] for r in res if r["status"] == 0)))
