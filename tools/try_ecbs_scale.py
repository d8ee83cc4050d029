"""ECBS w=1.3 lock-step batch throughput versus batch size (config C3 instances as in bench.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libmultirobotplanning_b200 as pkg
pkg.capi.init(0)
s32 = pkg.instances.load_set(os.path.join(ROOT, "tests/golden/bench_32x32.npz"))
insts = [i for i in s32 if i.n_agents == 100]
insts += [pkg.instances.synthetic_c3(b, k, 100) for k, b in enumerate(i for i in s32 if i.n_agents != 100)]
for n in [int(a) for a in sys.argv[1:]] or [100, 400, 800]:
    t = time.time()
    res = pkg.solver.solve_batch(pkg.solver.ECBS, insts[:n], w=1.3, max_hl=2000, max_seconds=300)
    dt = time.time() - t
    ok = sum(r["status"] == 0 for r in res)
    print("batch %4d: %d solved in %.2f s = %.1f inst/s; hl %d ll %d" % (
        n, ok, dt, ok / dt, sum(r["hl_expanded"] for r in res), sum(r["ll_expanded"] for r in res)), flush=True)
