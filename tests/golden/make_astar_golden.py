"""Low-level replans from the reference's OWN AStar::search (include/libMultiRobotPlanning/a_star.hpp:63-161)
driven through its own Environments (example/cbs.cpp, example/cbs_ta.cpp included unmodified into
oracle/_ref/astar_probe_cbs / astar_probe_cbs_ta by oracle/ref_build/Makefile) on seeded jobs ->
tests/golden/astar_probe_golden.json (cost and number of states per job; costs are unique, paths are not).

    make -C oracle/ref_build && python tests/golden/make_astar_golden.py

Jobs: a benchmark 32x32 map and an 8x8 map; random free start / goal pairs; vertex and edge constraints
taken from the agent's own unconstrained shortest path (so they force waits and detours), constraints on
the goal cell after arrival (the goal test must wait for them, example/cbs.cpp:266-276,287-291), and for the
cbs_ta Environment agents without a task (cbs_ta.cpp:283-319).  The jobs are regenerated from their seed by
tests/ (jobs() below).  Only runs where /root/reference exists."""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.path.join(ROOT, "oracle", "_ref")


def jobs(orc, s8, s32):
    """[(dimx, dimy, obstacles, [job...])], job = (variant, start_cell, goal_cell, vc[(t, cell)], ec[(t, from, to)])"""
    rng = np.random.default_rng(4242)
    out = []
    for inst, n_jobs in ((s32[0], 60), (s32[500], 40), (s8[1234], 40)):
        dx, dy = inst.dimx, inst.dimy
        free = np.ones((dy, dx), bool)
        free[inst.obstacles[:, 1], inst.obstacles[:, 0]] = False
        cells = np.flatnonzero(free.ravel())
        lst = []
        while len(lst) < n_jobs:
            variant = int(rng.integers(0, 2))
            s, g = (int(c) for c in rng.choice(cells, 2, replace=False))
            base = orc.lowlevel(dx, dy, inst.obstacles, variant, s, g)
            if base["status"] != 0:
                continue
            path = base["path"]  # (time, cell, g)
            vc, ec = [], []
            L = len(path)
            for _ in range(int(rng.integers(0, 7))):
                k = int(rng.integers(1, L)) if L > 1 else 0
                if k > 0:
                    vc.append((int(path[k][0]), int(path[k][1])))
            for _ in range(int(rng.integers(0, 3))):
                if L > 2:
                    k = int(rng.integers(0, L - 1))
                    ec.append((int(path[k][0]), int(path[k][1]), int(path[k + 1][1])))
            if rng.random() < 0.3:  # somebody crosses the goal after the agent got there
                vc.append((int(path[-1][0]) + int(rng.integers(1, 6)), g))
            goal = g
            if variant == 1 and rng.random() < 0.15:  # cbs_ta: an agent without a task only dodges
                goal = -1
                vc = [(int(rng.integers(1, 8)), s)] + [v for v in vc if v[1] != s][:2]
                ec = []
            r = orc.lowlevel(dx, dy, inst.obstacles, variant, s, goal, vc, ec, max_expanded=200000)
            if r["status"] != 0:  # the reference's search has no cap: only solvable jobs
                continue
            lst.append((variant, s, goal, sorted(set(vc)), sorted(set(ec))))
        out.append((dx, dy, inst.obstacles, lst))
    return out


def probe(dx, dy, obst, lst, variant):
    sel = [j for j in lst if j[0] == variant]
    txt = ["%d %d %d" % (dx, dy, len(obst))] + ["%d %d" % (x, y) for x, y in obst] + ["%d" % len(sel)]
    xy = lambda c: (c % dx, c // dx)
    for _, s, g, vc, ec in sel:
        gx, gy = xy(g) if g >= 0 else (-1, -1)
        txt.append("%d %d %d %d %d" % (*xy(s), gx, gy, len(vc)))
        txt += ["%d %d %d" % (t, *xy(c)) for t, c in vc]
        txt.append("%d" % len(ec))
        txt += ["%d %d %d %d %d" % (t, *xy(a), *xy(b)) for t, a, b in ec]
    exe = os.path.join(REF, "astar_probe_cbs" if variant == 0 else "astar_probe_cbs_ta")
    raw = subprocess.run([exe], input="\n".join(txt).encode(), stdout=subprocess.PIPE, check=True, cwd="/tmp",
                         timeout=600).stdout.decode().split("\n")
    res = [[int(v) for v in l.split()[1:3]] for l in raw if l.startswith("J")]
    assert len(res) == len(sel)
    it = iter(res)
    return [next(it) if j[0] == variant else None for j in lst]


if __name__ == "__main__":
    from libmultirobotplanning_b200 import instances as I
    from oracle import orc
    orc.build()
    s8 = I.load_set(os.path.join(HERE, "bench_8x8.npz"))
    s32 = I.load_set(os.path.join(HERE, "bench_32x32.npz"))
    g = []
    for dx, dy, obst, lst in jobs(orc, s8, s32):
        a, b = probe(dx, dy, obst, lst, 0), probe(dx, dy, obst, lst, 1)
        g.append([x if x is not None else y for x, y in zip(a, b)])
    with open(os.path.join(HERE, "astar_probe_golden.json"), "w") as f:
        json.dump(g, f)
    print([len(x) for x in g], "jobs")
