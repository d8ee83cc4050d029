import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libmultirobotplanning_b200 as pkg
pkg.capi.init(0)
s8 = pkg.instances.load_set(os.path.join(ROOT, "tests/golden/bench_8x8.npz"))
cap = int(sys.argv[1]) if len(sys.argv) > 1 else 500
t = time.time()
res = pkg.solver.solve_batch(pkg.solver.CBS, s8, max_hl=cap, max_seconds=300)
dt = time.time() - t
print("GPU CBS 8x8 full set cap %d: %d/2000 solved in %.2fs, hl total %d (%.0f HL/s)" % (
    cap, sum(r["status"] == 0 for r in res), dt, sum(r["hl_expanded"] for r in res),
    sum(r["hl_expanded"] for r in res) / dt))
