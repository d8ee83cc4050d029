"""The instances of the C3 ECBS batch that end capped: how far does a larger replan cap get?
usage: python tools/try_unsolved.py [max_ll ...]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import libmultirobotplanning_b200 as pkg  # noqa: E402

pkg.capi.init(0)
s32 = pkg.instances.load_set(os.path.join(ROOT, "tests", "golden", "bench_32x32.npz"))
insts, _ = bench.c3_shard(pkg, s32, 0)
res = pkg.solver.solve_batch(pkg.solver.ECBS, insts, w=1.3, max_hl=2000, max_seconds=120)
bad = [i for i, r in zip(insts, res) if r["status"] != 0]
print(len(bad), "capped at the default replan cap")
for cap in [int(a) for a in sys.argv[1:]] or [12000]:
    t0 = time.perf_counter()
    r2 = pkg.solver.solve_batch(pkg.solver.ECBS, bad, w=1.3, max_hl=2000, max_ll=cap, max_seconds=120)
    dt = time.perf_counter() - t0
    print("max_ll %d: %d/%d solved in %.2f s; hl %s; ll %s" % (
        cap, sum(r["status"] == 0 for r in r2), len(bad), dt,
        [r["hl_expanded"] for r in r2], [r["ll_expanded"] for r in r2]))
