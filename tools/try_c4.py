import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libmultirobotplanning_b200 as pkg
pkg.capi.init(0)
s32 = pkg.instances.load_set(os.path.join(ROOT, "tests/golden/bench_32x32.npz"))
for n in (20, 50, 100):
    base = [i for i in s32 if i.n_agents == n][0]
    inst = base.with_all_goals_potential()
    t = time.time()
    r = pkg.solver.solve_batch(pkg.solver.CBS_TA, [inst], max_hl=2000, max_seconds=60)[0]
    print(n, "agents: status", r["status"], "cost", r["cost"], "hl", r["hl_expanded"], "nTA", r["n_task_assignments"], "%.2fs" % (time.time() - t))
