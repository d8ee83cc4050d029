"""Config C3 at its stated sizes: the obstacle layouts of the 100 files
map_32by32_obst204_agents100_ex{k}, scaled to N agents by instances.synthetic_c3, ECBS w = 1.3.
Reports solved / cost over lower bound / validity (validate.validate_paths) per N.
usage: python tools/try_c3_scale.py [N ...] [--batch B] [--seconds S]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import libmultirobotplanning_b200 as pkg
pkg.capi.init(0)
args = [a for a in sys.argv[1:] if not a.startswith("--")]
opt = dict(a[2:].split("=") for a in sys.argv[1:] if a.startswith("--"))
B = int(opt.get("batch", 100)); S = float(opt.get("seconds", 120)); HL = int(opt.get("hl", 2000)); LL = int(opt.get("ll", 8000))
s32 = pkg.instances.load_set(os.path.join(ROOT, "tests/golden/bench_32x32.npz"))
base = [i for i in s32 if i.n_agents == 100][:B]
for n in [int(a) for a in args] or [120, 160, 200]:
    insts = [pkg.instances.synthetic_c3(b, k, n) for k, b in enumerate(base)]
    t = time.time()
    res = pkg.solver.solve_batch(pkg.solver.ECBS, insts, w=1.3, max_hl=HL, max_ll=LL, max_seconds=S)
    dt = time.time() - t
    ok = [(i, r) for i, r in zip(insts, res) if r["status"] == 0]
    bad = [(i.name, pkg.validate.validate_paths(i, r["paths"], 0)) for i, r in ok]
    bad = [b for b in bad if b[1]]
    ratio = max((r["cost"] / r["lower_bound"] for _, r in ok), default=0)
    st = np.bincount([r["status"] for r in res], minlength=3)
    print("N=%3d: %d/%d solved (status counts %s) in %.2f s = %.1f inst/s; max cost/LB %.4f; invalid %d %s; hl %d ll %d" % (
        n, len(ok), len(insts), st.tolist(), dt, len(ok) / dt, ratio, len(bad), bad[:2],
        sum(r["hl_expanded"] for r in res), sum(r["ll_expanded"] for r in res)), flush=True)
