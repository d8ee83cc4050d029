"""GPU parity: batched low-level replans (A* / A*-epsilon) vs the CPU oracle.
A*: cost bit-exact (optimal cost is unique).  A*-epsilon: tie-breaking is not
pinned by the reference, so the check is the contract of the algorithm: a valid
path with cost <= w * fmin and fmin <= optimal cost."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_map(rng, dimx, dimy, density):
    blocked = rng.random((dimy, dimx)) < density
    ys, xs = np.nonzero(blocked)
    return np.stack([xs, ys], 1).astype(np.int32), ~blocked


def _check_path(r, dimx, free, start, goal, vc, ec, variant):
    cells, g = r["cells"], r["g"]
    assert cells[0] == start and g[0] == 0
    vcs = {tuple(x) for x in vc}
    ecs = {tuple(x) for x in ec}
    for t in range(1, len(cells)):
        a, b = int(cells[t - 1]), int(cells[t])
        ax, ay, bx, by = a % dimx, a // dimx, b % dimx, b // dimx
        assert abs(ax - bx) + abs(ay - by) <= 1
        assert free[by, bx]
        assert (t, b) not in vcs and (t - 1, a, b) not in ecs
        step = 1
        if variant == 1 and a == b and (goal < 0 or a == goal):
            step = 0
        assert g[t] - g[t - 1] == step
    last_goal = max([t for t, c in vc if goal < 0 or c == goal], default=-1)
    assert goal < 0 or cells[-1] == goal
    assert len(cells) - 1 > last_goal
    assert g[-1] == r["cost"]


@pytest.mark.parametrize("dims", [(8, 8), (32, 32), (50, 40)])
@pytest.mark.parametrize("variant", [0, 1])
def test_astar_costs_match_oracle(capi, orc, dims, variant):
    dimx, dimy = dims
    rng = np.random.default_rng(dimx * 100 + variant)
    obst, free = _rand_map(rng, dimx, dimy, 0.2)
    fc = np.flatnonzero(free.ravel())
    m = capi.Map(dimx, dimy, obst)
    jobs, ref = [], []
    goals = rng.choice(fc, 6, replace=False)
    gxy = np.stack([goals % dimx, goals // dimx], 1)
    fields = capi.bfs_fields(dimx, dimy, obst, gxy)
    for k in range(60):
        gi = int(rng.integers(0, len(goals)))
        goal = int(goals[gi])
        reach = np.flatnonzero(fields[gi] != capi.INF)
        start = int(rng.choice(reach))
        # constraints along and around a shortest path so that they bite
        nv, ne = int(rng.integers(0, 6)), int(rng.integers(0, 4))
        vc = [(int(rng.integers(1, 25)), int(rng.choice(reach))) for _ in range(nv)]
        if k % 3 == 0:
            vc.append((int(fields[gi][start]) + int(rng.integers(0, 4)), goal))
        ec = []
        for _ in range(ne):
            a = int(rng.choice(reach))
            ax, ay = a % dimx, a // dimx
            nb = [(ax + dx, ay + dy) for dx, dy in ((1, 0), (-1, 0), (0, 1), (0, -1), (0, 0))]
            nb = [x + dimx * y for x, y in nb if 0 <= x < dimx and 0 <= y < dimy and free[y, x]]
            ec.append((int(rng.integers(0, 20)), a, int(rng.choice(nb))))
        use_field = variant == 1 or k % 2 == 0
        no_task = variant == 1 and k % 10 == 9
        jobs.append({"map": 0, "start": start, "goal": -1 if no_task else goal,
                     "field": -1 if (no_task or not use_field) else gi, "vc": vc, "ec": ec})
        ref.append(orc.lowlevel(dimx, dimy, obst, variant, start, -1 if no_task else goal,
                                vc=vc, ec=ec, w=0.0, max_expanded=200000))
    got = capi.lowlevel_batch([m], fields, jobs, variant=variant, w=0.0,
                              max_expanded=8000, path_cap=512)
    n_ok = 0
    for j, r, o in zip(jobs, got, ref):
        if o["status"] != 0:
            assert r["status"] != 0
            continue
        assert r["status"] == 0, (j, r, o["cost"])
        assert r["cost"] == o["cost"], (j, r["cost"], o["cost"])
        assert r["fmin"] == r["cost"]
        _check_path(r, dimx, free, j["start"], j["goal"], j["vc"], j["ec"], variant)
        n_ok += 1
    assert n_ok >= 50
    m.close()


def test_astar_eps_bound_and_focal(capi, orc, set32):
    inst = next(i for i in set32 if i.name == "map_32by32_obst204_agents30_ex0")
    dimx = 32
    free = np.ones((32, 32), bool)
    free[inst.obstacles[:, 1], inst.obstacles[:, 0]] = False
    fields = capi.bfs_fields(32, 32, inst.obstacles, inst.goals)
    m = capi.Map(32, 32, inst.obstacles)
    N = inst.n_agents
    starts = inst.cell(inst.starts)
    goals = inst.cell(inst.goals)
    # other agents' paths: optimal single-agent paths (A*)
    base = capi.lowlevel_batch([m], fields, [{"start": int(starts[i]), "goal": int(goals[i]),
                                             "field": i} for i in range(N)])
    T = max(len(r["cells"]) for r in base)
    table = np.zeros((1, N, T), np.int32)
    tlen = np.zeros((1, N), np.int32)
    for i, r in enumerate(base):
        table[0, i, :len(r["cells"])] = r["cells"]
        tlen[0, i] = len(r["cells"])
    for w in (1.0, 1.3, 2.0):
        jobs = [{"start": int(starts[i]), "goal": int(goals[i]), "field": i, "table": 0,
                 "self": i, "vc": [(5, int(base[i]["cells"][min(5, len(base[i]["cells"]) - 1)]))]}
                for i in range(N)]
        got = capi.lowlevel_batch([m], fields, jobs, variant=0, w=w, max_expanded=8000,
                                  path_cap=512, tables=table, table_len=tlen)
        for i, (j, r) in enumerate(zip(jobs, got)):
            opt = orc.lowlevel(32, 32, inst.obstacles, 0, j["start"], j["goal"], vc=j["vc"])
            assert r["status"] == 0 and opt["status"] == 0
            assert r["fmin"] <= opt["cost"]
            assert np.float32(r["cost"]) <= np.float32(r["fmin"]) * np.float32(w)
            assert r["cost"] >= opt["cost"]
            if w == 1.0:
                assert r["cost"] == opt["cost"]
            _check_path(r, dimx, free, j["start"], j["goal"], j["vc"], [], 0)
    # the focal heuristic must actually steer: with w = 2 the number of
    # conflicts against the table must not exceed that of the A* path
    def conflicts(cells, self_idx):
        n = 0
        for t, c in enumerate(cells):
            for a in range(N):
                if a == self_idx:
                    continue
                pa = table[0, a, min(t, tlen[0, a] - 1)]
                n += int(pa == c)
        return n
    worse = sum(conflicts(got[i]["cells"], i) > conflicts(base[i]["cells"], i) for i in range(N))
    assert worse <= N // 4
    m.close()


def test_lowlevel_edge_cases(capi):
    m = capi.Map(5, 1, [])
    f = capi.bfs_fields(5, 1, [], [[4, 0]])
    r = capi.lowlevel_batch([m], f, [
        {"start": 0, "goal": 4, "field": 0},
        {"start": 0, "goal": 4, "field": 0, "vc": [(4, 4)]},
        {"start": 0, "goal": 4, "field": 0, "vc": [(9, 4)]},
        {"start": 0, "goal": 4, "field": 0, "ec": [(0, 0, 1)]},
        {"start": 4, "goal": 4, "field": 0},
        {"start": 0, "goal": 4, "field": -1},
    ])
    assert [x["cost"] for x in r] == [4, 5, 10, 5, 0, 4]
    assert list(r[4]["cells"]) == [4]
    # cbs_ta: waiting on the goal is free
    r = capi.lowlevel_batch([m], f, [{"start": 0, "goal": 4, "field": 0, "vc": [(9, 4)]}],
                            variant=1)
    assert r[0]["cost"] == 6
    # walled-in start: no solution (OPEN runs empty)
    m2 = capi.Map(3, 1, [[1, 0]])
    f2 = capi.bfs_fields(3, 1, [[1, 0]], [[2, 0]])
    r = capi.lowlevel_batch([m2], f2, [{"start": 0, "goal": 2, "field": 0}])
    assert r[0]["status"] == 1
    # expansion cap
    m3 = capi.Map(3, 1, [[1, 0]])
    r = capi.lowlevel_batch([m3], None, [{"start": 0, "goal": 2, "field": -1}],
                            max_expanded=50)
    assert r[0]["status"] == 2
    assert capi.lowlevel_batch([m], f, []) == []


def _same_results(a, b):
    assert len(a) == len(b)
    for k, (x, y) in enumerate(zip(a, b)):
        for key in ("status", "cost", "fmin", "expanded"):
            assert x[key] == y[key], (k, key, x[key], y[key])
        if x["status"] == 0:
            assert np.array_equal(x["cells"], y["cells"]), k
            assert np.array_equal(x["g"], y["g"]), k


@pytest.mark.parametrize("dims", [(8, 8), (32, 32), (20, 13), (32, 9)])
def test_tile_kernel_equals_general_kernel(capi, dims, monkeypatch):
    """Single-tile maps with cbs/ecbs moves run in the shared-memory kernel
    (lowlevel_tile.cu); MRP_LL_GENERIC=1 sends the same jobs through the general kernel.
    Same selection rule, tie-breaking and node numbering: status, cost, fmin, the number
    of expansions and every path must be identical, for A* and for A*-epsilon with the
    occupancy planes, with the exact focal pass only (MRP_LL_TILE_NOOCC), and when the
    visited bitmap is too short (MRP_LL_TILE_TB: those jobs are redone by the general
    kernel)."""
    dimx, dimy = dims
    rng = np.random.default_rng(dimx * 1000 + dimy)
    obst, free = _rand_map(rng, dimx, dimy, 0.15)
    fc = np.flatnonzero(free.ravel())
    m = capi.Map(dimx, dimy, obst)
    N = min(24, len(fc) // 3)
    pick = rng.choice(fc, 2 * N, replace=False)
    starts, goals = pick[:N], pick[N:]
    fields = capi.bfs_fields(dimx, dimy, obst, np.stack([goals % dimx, goals // dimx], 1))
    ok = [i for i in range(N) if fields[i][starts[i]] != capi.INF]
    base = capi.lowlevel_batch([m], fields, [{"start": int(starts[i]), "goal": int(goals[i]), "field": i}
                                             for i in range(N)])
    # two path tables: optimal paths, and the same with agents that have no path yet
    # (ECBS root construction) and two agents on one cell (>= 2 per cell in the planes)
    T = max([len(r["cells"]) for r in base if r["status"] == 0] + [2])
    table = np.zeros((2, N, T), np.int32)
    tlen = np.zeros((2, N), np.int32)
    for i, r in enumerate(base):
        if r["status"] != 0:
            continue
        for b in range(2):
            table[b, i, :len(r["cells"])] = r["cells"]
            tlen[b, i] = len(r["cells"])
    tlen[1, ok[0]] = 0
    if len(ok) > 3:
        table[1, ok[1]] = table[1, ok[2]]
        tlen[1, ok[1]] = tlen[1, ok[2]]
    jobs = []
    for k in range(160):
        i = int(rng.choice(ok))
        cells = base[i]["cells"]
        reach = np.flatnonzero(fields[i] != capi.INF)
        vc = [(int(rng.integers(1, len(cells) + 6)), int(rng.choice(reach))) for _ in range(int(rng.integers(0, 5)))]
        # constraints on the agent's own optimal path so that they bite, sometimes on its goal
        for _ in range(int(rng.integers(0, 4))):
            t = int(rng.integers(1, len(cells) + 1))
            vc.append((t, int(cells[min(t, len(cells) - 1)])))
        if k % 4 == 0:
            vc.append((len(cells) + int(rng.integers(0, 5)), int(goals[i])))
        ec = []
        for _ in range(int(rng.integers(0, 3))):
            t = int(rng.integers(0, max(1, len(cells) - 1)))
            ec.append((t, int(cells[min(t, len(cells) - 1)]), int(cells[min(t + 1, len(cells) - 1)])))
        if k == 7:  # more constraints than the shared-memory cache holds
            vc += [(int(rng.integers(40, 60)), int(rng.choice(reach))) for _ in range(120)]
        jobs.append({"start": int(starts[i]), "goal": int(goals[i]), "field": i if k % 5 else -1,
                     "table": k % 2, "self": i if k % 7 else -1, "vc": vc, "ec": ec})
    for w, cap in ((0.0, 8000), (1.0, 8000), (1.3, 8000), (2.0, 8000), (1.3, 40)):
        kw = dict(variant=0, w=w, max_expanded=cap, path_cap=256, tables=table, table_len=tlen)
        monkeypatch.setenv("MRP_LL_GENERIC", "1")
        ref = capi.lowlevel_batch([m], fields, jobs, **kw)
        monkeypatch.delenv("MRP_LL_GENERIC")
        _same_results(capi.lowlevel_batch([m], fields, jobs, **kw), ref)
        monkeypatch.setenv("MRP_LL_TILE_NOOCC", "1")
        _same_results(capi.lowlevel_batch([m], fields, jobs, **kw), ref)
        monkeypatch.delenv("MRP_LL_TILE_NOOCC")
        monkeypatch.setenv("MRP_LL_TILE_TB", "32")
        _same_results(capi.lowlevel_batch([m], fields, jobs, **kw), ref)
        monkeypatch.delenv("MRP_LL_TILE_TB")
        assert sum(r["status"] == 0 for r in ref) >= (100 if cap > 40 else 1)
    m.close()
