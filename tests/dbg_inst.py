import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import libmultirobotplanning_b200 as pkg
from oracle import orc
pkg.capi.init(0)
s32 = pkg.instances.load_set(os.path.join(ROOT, "tests/golden/bench_32x32.npz"))
i = next(x for x in s32 if x.name == sys.argv[1])
r = pkg.solver.solve_batch(pkg.solver.CBS, [i], max_hl=300000, max_seconds=100)[0]
print({k: v for k, v in r.items() if k != "paths"})
o = orc.cbs(i.dimx, i.dimy, i.obstacles, i.starts, i.goals, (300000, 0, 60.0))
print({k: v for k, v in o.items() if k != "paths"})
