"""Does a large ECBS batch slow down the CBS 8x8 batch that follows it in the same process?"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import libmultirobotplanning_b200 as pkg  # noqa: E402

pkg.capi.init(0)
g = os.path.join(ROOT, "tests", "golden")
s32 = pkg.instances.load_set(os.path.join(g, "bench_32x32.npz"))
s8 = pkg.instances.load_set(os.path.join(g, "bench_8x8.npz"))


def cbs(tag):
    pkg.solver.solve_batch(pkg.solver.CBS, s8[:4], max_hl=50)
    for _ in range(2):
        t0 = time.perf_counter()
        res = pkg.solver.solve_batch(pkg.solver.CBS, s8, max_hl=500, max_seconds=120)
        print(tag, "CBS 8x8: %.2f s" % (time.perf_counter() - t0), flush=True)


cbs("fresh")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
insts = []
for r in range((n + 999) // 1000):
    insts += bench.c3_shard(pkg, s32, r if r == 0 else 1000 + r)[0]
t0 = time.perf_counter()
pkg.solver.solve_batch(pkg.solver.ECBS, insts[:n], w=1.3, max_hl=2000, max_seconds=120)
print("ECBS %d: %.2f s" % (n, time.perf_counter() - t0), flush=True)
cbs("after")
