"""GPU parity of the end-to-end searches (batched drivers + drop-in binaries):
optimal sums-of-costs of CBS / CBS-TA bit-exact against the reference's pinned
answers and the oracle; ECBS within the w bound."""
import os
import re
import subprocess

import numpy as np
import pytest
import yaml

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")


class Inst:
    def __init__(self, d):
        self.dimx, self.dimy = d["dimx"], d["dimy"]
        self.obstacles = np.array(d["obstacles"], np.int32).reshape(-1, 2)
        self.starts = np.array(d["starts"], np.int32).reshape(-1, 2)
        if "goals" in d:
            self.goals = np.array(d["goals"], np.int32).reshape(-1, 2)
        else:
            self.potential_goals = [np.array(g, np.int32).reshape(-1, 2)
                                    for g in d["potentialGoals"]]

    def cell(self, xy):
        xy = np.asarray(xy)
        return xy[..., 0] + self.dimx * xy[..., 1]


def check_solution(inst, paths, mode):
    """Validity under the reference's own conflict semantics (same loop bound)."""
    free = np.ones((inst.dimy, inst.dimx), bool)
    if len(inst.obstacles):
        free[inst.obstacles[:, 1], inst.obstacles[:, 0]] = False
    for a, p in enumerate(paths):
        assert tuple(p[0][:2]) == tuple(inst.starts[a]) and p[0][2] == 0
        for t in range(len(p)):
            assert free[p[t][1], p[t][0]]
            if t:
                assert abs(p[t][0] - p[t - 1][0]) + abs(p[t][1] - p[t - 1][1]) <= 1
    T = max(len(p) for p in paths) - (1 if mode == 0 else 0)
    pos = lambda a, t: tuple(paths[a][min(t, len(paths[a]) - 1)][:2])
    for t in range(T):
        for a in range(len(paths)):
            for b in range(a + 1, len(paths)):
                assert pos(a, t) != pos(b, t)
                assert not (pos(a, t) == pos(b, t + 1) and pos(a, t + 1) == pos(b, t))


def test_reference_fixtures_batched(capi, ref_fixtures):
    from libmultirobotplanning_b200 import solver
    for name, d in ref_fixtures.items():
        exp = d["expected"]
        inst = Inst(d)
        if "cbs_cost" in exp:
            r = solver.solve_batch(solver.CBS, [inst], max_hl=5000)[0]
            assert r["status"] == 0 and r["cost"] == exp["cbs_cost"], name
            check_solution(inst, r["paths"], 0)
            r = solver.solve_batch(solver.ECBS, [inst], w=1.0, max_hl=5000)[0]
            assert r["status"] == 0 and r["cost"] == exp["ecbs_w1_cost"], name
        if "cbs_ta_cost" in exp:
            r = solver.solve_batch(solver.CBS_TA, [inst], max_hl=5000)[0]
            assert r["status"] == 0 and r["cost"] == exp["cbs_ta_cost"], name
            if "agent0_last" in exp:
                x, y, g = r["paths"][0][-1]
                assert {"x": x, "y": y, "t": g} == exp["agent0_last"]
            if "agent1_last_xy" in exp:
                assert list(r["paths"][1][-1][:2]) == exp["agent1_last_xy"]


def test_swap_fixtures(capi, ref_fixtures):
    from libmultirobotplanning_b200 import solver
    for name, cost, mk in (("mapf_swap2", 12, 6), ("mapf_swap4", 28, 8), ("mapf_simple1b", 8, 4)):
        r = solver.solve_batch(solver.CBS, [Inst(ref_fixtures[name])], max_hl=5000)[0]
        assert (r["status"], r["cost"], r["makespan"]) == (0, cost, mk), name


def test_cbs_8x8_set_matches_oracle_golden(capi, set8, oracle_golden):
    """Config C2: CBS over the 8x8 benchmark set, all instances in one lock-step
    batch; sums of costs bit-exact on every instance of the golden file."""
    from libmultirobotplanning_b200 import solver
    names = [n for n in oracle_golden["cbs"] if n.startswith("map_8by8")]
    by = {i.name: i for i in set8}
    insts = [by[n] for n in names]
    res = solver.solve_batch(solver.CBS, insts, max_hl=20000, max_seconds=240)
    sums = {}
    for n, i, r in zip(names, insts, res):
        g = oracle_golden["cbs"][n]
        if g["status"] != 0 or r["status"] != 0:
            continue  # capped on one side: nothing to compare
        assert r["cost"] == g["cost"], n
        k = i.n_agents
        sums[k] = sums.get(k, 0) + r["cost"]
    assert [sums[k] for k in (1, 2, 3, 4, 5)] == [599, 1167, 1765, 2418, 2979]
    for i, r in list(zip(insts, res))[::23]:
        if r["status"] == 0:
            check_solution(i, r["paths"], 0)


def test_cbs_matches_reference_binaries(capi, set8, set32):
    """582 optimal sums-of-costs produced by the UNMODIFIED reference cbs binary
    (tests/golden/ref_binary_golden.json, see make_ref_golden.py)."""
    import json
    from libmultirobotplanning_b200 import solver
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_binary_golden.json")))
    by = {i.name: i for i in set8 + set32}
    for dims in ((8, 8), (32, 32)):
        names = [n for n in sorted(g["cbs"]) if (by[n].dimx, by[n].dimy) == dims]
        res = solver.solve_batch(solver.CBS, [by[n] for n in names], max_hl=300000,
                                 max_seconds=400)
        for n, r in zip(names, res):
            assert r["status"] == 0, n
            assert r["cost"] == g["cbs"][n]["cost"], n
    # ECBS w = 1.3: within the bound, and in the neighbourhood of the reference's cost
    names = sorted(g["ecbs_w1.3"])
    res = solver.solve_batch(solver.ECBS, [by[n] for n in names], w=1.3, max_hl=20000,
                             max_seconds=300)
    for n, r in zip(names, res):
        assert r["status"] == 0, n
        assert np.float32(r["cost"]) <= np.float32(r["lower_bound"]) * np.float32(1.3)
        assert r["cost"] <= 1.05 * g["ecbs_w1.3"][n]["cost"], n


def test_cbs_32x32_and_ecbs_bound(capi, set32, oracle_golden):
    from libmultirobotplanning_b200 import solver
    names = [n for n in oracle_golden["cbs"] if n.startswith("map_32by32")]
    by = {i.name: i for i in set32}
    insts = [by[n] for n in names]
    res = solver.solve_batch(solver.CBS, insts, max_hl=3000, max_seconds=240)
    n_cmp = 0
    for n, r in zip(names, res):
        g = oracle_golden["cbs"][n]
        if g["status"] == 0 and r["status"] == 0:
            assert (r["cost"]) == (g["cost"]), n
            n_cmp += 1
    assert n_cmp >= 15
    # config C1: ECBS w=1.3 on agents10_ex1; optimal SOC 236 => cost in [236, 306]
    i = by["map_32by32_obst204_agents10_ex1"]
    for w in (1.0, 1.3):
        r = solver.solve_batch(solver.ECBS, [i], w=w, max_hl=3000)[0]
        assert r["status"] == 0 and 236 <= r["cost"] <= 306
        assert np.float32(r["cost"]) <= np.float32(r["lower_bound"]) * np.float32(w)
        assert r["lower_bound"] <= 236
        if w == 1.0:
            assert r["cost"] == 236
        check_solution(i, r["paths"], 0)
    # ECBS on crowded instances (C3 regime): 50 agents, w = 1.3
    crowd = [x for x in set32 if x.n_agents == 50][:8]
    res = solver.solve_batch(solver.ECBS, crowd, w=1.3, max_hl=2000, max_seconds=120)
    solved = 0
    for i, r in zip(crowd, res):
        if r["status"] == 0:
            solved += 1
            assert np.float32(r["cost"]) <= np.float32(r["lower_bound"]) * np.float32(1.3)
            check_solution(i, r["paths"], 0)
    assert solved >= 6


def test_cbs_ta_c4_like(capi, orc, set32):
    """Config C4 shape at a size the oracle finishes quickly: every agent may
    take any goal of the instance; optimal cost equals the oracle's."""
    from libmultirobotplanning_b200 import solver
    base = next(i for i in set32 if i.name == "map_32by32_obst204_agents10_ex3")
    inst = base.with_all_goals_potential()
    r = solver.solve_batch(solver.CBS_TA, [inst], max_hl=3000)[0]
    o = orc.cbs_ta(inst.dimx, inst.dimy, inst.obstacles, inst.starts,
                   [g.tolist() for g in inst.potential_goals], caps=(3000, 0, 60.0))
    assert r["status"] == 0 and o["status"] == 0
    assert r["cost"] == o["cost"]
    check_solution(inst, r["paths"], 1)


def test_ecbs_ta(capi, orc, ref_fixtures, set32):
    """ecbs_ta: the reference pins the cbs_ta answers at
    w = 1.0 (test/test_ecbs_ta.py:25-39); beyond that the w bound must hold."""
    from libmultirobotplanning_b200 import solver
    for name, d in ref_fixtures.items():
        exp = d["expected"]
        if "cbs_ta_cost" not in exp:
            continue
        inst = Inst(d)
        r = solver.solve_batch(solver.ECBS_TA, [inst], w=1.0, max_hl=5000)[0]
        assert r["status"] == 0 and r["cost"] == exp["cbs_ta_cost"], name
        if "agent0_last" in exp:
            x, y, g = r["paths"][0][-1]
            assert {"x": x, "y": y, "t": g} == exp["agent0_last"]
        if "agent1_last_xy" in exp:
            assert list(r["paths"][1][-1][:2]) == exp["agent1_last_xy"]
    base = next(i for i in set32 if i.name == "map_32by32_obst204_agents10_ex3")
    inst = base.with_all_goals_potential()
    opt = solver.solve_batch(solver.CBS_TA, [inst], max_hl=3000)[0]
    r1 = solver.solve_batch(solver.ECBS_TA, [inst], w=1.0, max_hl=3000)[0]
    o1 = orc.ecbs_ta(inst.dimx, inst.dimy, inst.obstacles, inst.starts,
                     [g.tolist() for g in inst.potential_goals], 1.0, caps=(3000, 0, 60.0))
    assert r1["status"] == 0 and o1["status"] == 0 and opt["status"] == 0
    assert r1["cost"] == o1["cost"] == opt["cost"]
    r = solver.solve_batch(solver.ECBS_TA, [inst], w=1.3, max_hl=3000)[0]
    assert r["status"] == 0
    assert np.float32(r["cost"]) <= np.float32(r["lower_bound"]) * np.float32(1.3)
    assert opt["cost"] <= r["cost"] <= 1.3 * opt["cost"]
    check_solution(inst, r["paths"], 1)


def test_cli_binaries(capi, ref_fixtures, tmp_path):
    from libmultirobotplanning_b200 import instances as I
    for name, tool, extra, key in (("mapf_simple1", "cbs", [], "cbs_cost"),
                                   ("mapf_circle", "cbs", [], "cbs_cost"),
                                   ("mapf_atGoal", "ecbs", ["-w", "1.0"], "ecbs_w1_cost"),
                                   ("mapf_simple1", "ecbs", ["--suboptimality=1.0"], "ecbs_w1_cost"),
                                   ("mapfta_simple1_a2", "cbs_ta", [], "cbs_ta_cost"),
                                   ("mapfta_simple1_a3", "cbs_ta", ["--maxTaskAssignments", "5"],
                                    "cbs_ta_cost"),
                                   ("mapfta_simple1_a2", "ecbs_ta", ["-w", "1.0"], "cbs_ta_cost"),
                                   ("mapfta_simple1_a1", "ecbs_ta", [], "cbs_ta_cost")):
        d = ref_fixtures[name]
        x = Inst(d)
        inst = I.Instance(name, x.dimx, x.dimy, x.obstacles, x.starts,
                          getattr(x, "goals", None), getattr(x, "potential_goals", None))
        inp, out = str(tmp_path / "in.yaml"), str(tmp_path / "out.yaml")
        if os.path.exists(out):
            os.remove(out)
        I.save_yaml(inst, inp)
        r = subprocess.run([os.path.join(BIN, tool), "-i", inp, "-o", out] + extra,
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert "Planning successful!" in r.stdout and "done; cost:" in r.stdout
        text = open(out).read()
        y = yaml.safe_load(text)
        assert y["statistics"]["cost"] == d["expected"][key]
        assert list(y["statistics"].keys())[:5] == ["cost", "makespan", "runtime",
                                                    "highLevelExpanded", "lowLevelExpanded"]
        assert ("numTaskAssignments" in y["statistics"]) == (tool in ("cbs_ta", "ecbs_ta"))
        assert re.search(r"^schedule:\n  agent0:\n    - x: \d+\n      y: \d+\n      t: 0\n", text, re.M)
        assert sorted(y["schedule"]) == ["agent%d" % a for a in range(len(x.starts))]
        if "agent0_last" in d["expected"]:
            assert y["schedule"]["agent0"][-1] == d["expected"]["agent0_last"]


def test_environment_adapter(capi):
    from libmultirobotplanning_b200 import solver
    # two agents walking into each other on a 3x1 corridor: vertex conflict at t=1
    assert solver.lib().mrph_environment_selftest() == 10


def test_reference_templates_drive_gpu_environment(capi, ref_fixtures, tmp_path):
    """oracle/_ref/templates_gpuenv = the reference's UNMODIFIED CBS / ECBS /
    CBSTA templates instantiated with this repository's GPU-backed Environment
    (built by oracle/ref_build where /root/reference exists; the binary travels
    to the GPU box).  The reference's own loops must reach its pinned answers
    through our callbacks."""
    from libmultirobotplanning_b200 import instances as I
    exe = os.path.join(ROOT, "oracle", "_ref", "templates_gpuenv")
    # built by __graft_entry__.build() in the container that holds /root/reference and shipped to
    # the GPU box with the snapshot (git-ignored, not gpurun-ignored): its absence is a failure
    assert os.path.exists(exe), "oracle/_ref/templates_gpuenv is missing: run __graft_entry__.build() " \
                                "where /root/reference exists before going to the GPU box"
    for name, d in ref_fixtures.items():
        exp = d["expected"]
        x = Inst(d)
        inst = I.Instance(name, x.dimx, x.dimy, x.obstacles, x.starts,
                          getattr(x, "goals", None), getattr(x, "potential_goals", None))
        inp = str(tmp_path / (name + ".yaml"))
        I.save_yaml(inst, inp)
        runs = []
        if "cbs_cost" in exp:
            runs = [("cbs", [], exp["cbs_cost"]), ("ecbs", ["1.0"], exp["ecbs_w1_cost"])]
        if "cbs_ta_cost" in exp:
            runs = [("cbs_ta", [], exp["cbs_ta_cost"])]
        for algo, extra, want in runs:
            r = subprocess.run([exe, algo, inp] + extra, capture_output=True, text=True)
            assert r.returncode == 0, (name, algo, r.stderr[-300:])
            assert ("cost %d" % want) in r.stdout, (name, algo, r.stdout[-200:])


def test_movingai_instance_cbs_and_ecbs(capi, orc, tmp_path):
    """A movingai-format map (non-square, larger than one 32x32 tile, so the distance fields
    come from the queue BFS kernel) through the ingestion path: CBS cost equals the oracle's,
    ECBS stays within w."""
    from libmultirobotplanning_b200 import instances as I
    from libmultirobotplanning_b200 import solver
    rng = np.random.default_rng(5)
    W, H = 48, 40
    grid = np.where(rng.random((H, W)) < 0.15, "@", ".")
    grid[rng.random((H, W)) < 0.03] = "T"
    free = [(x, y) for y in range(H) for x in range(W) if grid[y, x] == "."]
    (tmp_path / "m.map").write_text("type octile\nheight %d\nwidth %d\nmap\n" % (H, W) +
                                    "\n".join("".join(r) for r in grid) + "\n")
    # starts / goals inside the largest component so that every agent has a path
    _, _, obst = I.load_movingai_map(str(tmp_path / "m.map"))
    f = orc.bfs_fields(W, H, obst, [free[0]])[0].reshape(H, W)
    comp = [c for c in free if f[c[1], c[0]] != capi.INF]
    pick = rng.choice(len(comp), 12, replace=False)
    lines = ["version 1"]
    for k in range(6):
        s, g = comp[pick[2 * k]], comp[pick[2 * k + 1]]
        lines.append("\t".join(map(str, [k // 2, "m.map", W, H, s[0], s[1], g[0], g[1], 1.0])))
    (tmp_path / "m.scen").write_text("\n".join(lines) + "\n")
    inst = I.movingai_instances(str(tmp_path / "m.scen"), str(tmp_path / "m.map"), 6, 1)[0]
    want = orc.cbs(W, H, inst.obstacles, inst.starts, inst.goals)
    assert want["status"] == 0
    r = solver.solve_batch(solver.CBS, [inst], max_hl=20000)[0]
    assert r["status"] == 0 and r["cost"] == want["cost"]
    check_solution(inst, r["paths"], 0)
    e = solver.solve_batch(solver.ECBS, [inst], w=1.2, max_hl=20000)[0]
    assert e["status"] == 0 and want["cost"] <= e["cost"] <= np.float32(1.2) * np.float32(want["cost"])


def test_lanes_give_the_same_answers(capi, set8, monkeypatch):
    """Sub-batches in separate lanes (own streams, staging buffers and replan
    workspace; mrp_set_lane) must not change a single result: instances do not
    interact.  Same batch with one lane and with several."""
    from libmultirobotplanning_b200 import solver
    insts = [i for i in set8 if i.n_agents in (4, 6, 8)][:288]
    out = {}
    for lanes in ("1", "6"):
        monkeypatch.setenv("MRP_HOST_LANES", lanes)
        monkeypatch.setenv("MRP_HOST_LANE_SIZE", "16")
        res = solver.solve_batch(solver.CBS, insts, max_hl=300, max_seconds=120)
        out[lanes] = [(r["status"], r["cost"], r["makespan"], r["hl_expanded"], r["ll_expanded"]) for r in res]
    assert out["1"] == out["6"]
    assert sum(s == 0 for s, *_ in out["1"]) > 200


def test_lanes_concurrent_calls(capi):
    """Host threads in different lanes call the C ABI at the same time."""
    import threading
    rng = np.random.default_rng(5)
    blocked = rng.random((48, 40)) < 0.2
    ys, xs = np.nonzero(blocked)
    obst = np.stack([xs, ys], 1).astype(np.int32)
    goals = np.array([[x, y] for x, y in zip(rng.integers(0, 40, 24), rng.integers(0, 48, 24))], np.int32)
    want = capi.bfs_fields(40, 48, obst, goals)
    cell = rng.integers(0, 300, (64, 12)).astype(np.int32)
    ln = np.full(64, 12, np.int32)
    want_c = capi.count_conflicts(cell, ln, 0)
    errors = []

    def work(k):
        try:
            capi.set_lane(k)
            for _ in range(6):
                assert np.array_equal(capi.bfs_fields(40, 48, obst, goals), want)
                assert capi.count_conflicts(cell, ln, 0) == want_c
        except Exception as e:  # noqa: BLE001
            errors.append(repr(e))

    th = [threading.Thread(target=work, args=(k,)) for k in range(1, 6)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors


def test_sweep_binary_streams_yaml_files(capi, set8, tmp_path):
    """bin/mapf_sweep: the benchmark sweep as one streaming run (reader thread + batched driver).
    Same costs as the batch API on 60 files of the 8x8 set, output.yaml per solved file in the
    reference's layout (example/cbs.cpp:637-661)."""
    from libmultirobotplanning_b200 import instances as I
    from libmultirobotplanning_b200 import solver
    insts = [i for i in set8 if i.n_agents <= 6][:60]
    files = []
    for i in insts:
        p = str(tmp_path / (i.name + ".yaml"))
        I.save_yaml(i, p)
        files.append(p)
    lst = tmp_path / "files.txt"
    lst.write_text("\n".join(files) + "\n")
    outdir = tmp_path / "out"
    outdir.mkdir()
    csv = tmp_path / "res.csv"
    r = subprocess.run([os.path.join(BIN, "mapf_sweep"), "--algo", "cbs", "--batch", "16", "--maxHighLevelExpansions",
                        "2000", "--outputDir", str(outdir), "--csv", str(csv), "--list", str(lst)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-500:]
    rows = [l.split(",") for l in csv.read_text().strip().splitlines()[1:]]
    assert [row[0] for row in rows] == files
    want = solver.solve_batch(solver.CBS, insts, max_hl=2000)
    for row, w, i in zip(rows, want, insts):
        assert int(row[1]) == w["status"], i.name
        if w["status"] == 0:
            assert int(row[2]) == w["cost"] and int(row[3]) == w["makespan"], i.name
            y = yaml.safe_load((outdir / (i.name + ".output.yaml")).read_text())
            assert y["statistics"]["cost"] == w["cost"]
            assert sorted(y["schedule"]) == sorted("agent%d" % a for a in range(i.n_agents))


def test_output_yaml_replays_like_visualize_py(capi, set32, tmp_path):
    """The schedule our `cbs` writes, consumed the way the reference's example/visualize.py consumes it
    (matplotlib is not in this image, so its accesses are replayed): schedule["schedule"][agent name]
    for every agent of the input (visualize.py:57-63), T = max last t (:63), getState's linear
    interpolation between consecutive entries (:113-127) for the frames i/10 (:100-104), and its
    agent-agent collision test ||p1 - p2|| < 0.7 (:111-123), which must never fire on a CBS solution."""
    from libmultirobotplanning_b200 import instances as I
    inst = next(i for i in set32 if i.name == "map_32by32_obst204_agents10_ex1")
    inp, outp = str(tmp_path / "in.yaml"), str(tmp_path / "out.yaml")
    I.save_yaml(inst, inp)
    r = subprocess.run([os.path.join(BIN, "cbs"), "-i", inp, "-o", outp], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Planning successful!" in r.stdout
    with open(inp) as f:
        mp = yaml.safe_load(f)
    with open(outp) as f:
        schedule = yaml.safe_load(f)
    names = [d["name"] for d in mp["agents"]]
    T = 0
    for d in mp["agents"]:
        s = schedule["schedule"][d["name"]]            # KeyError here = visualize.py would crash
        assert (s[0]["x"], s[0]["y"], s[0]["t"]) == (d["start"][0], d["start"][1], 0)
        assert (s[-1]["x"], s[-1]["y"]) == tuple(d["goal"])
        assert all(b["t"] > a["t"] for a, b in zip(s, s[1:]))  # dt != 0 in getState
        T = max(T, s[-1]["t"])
    assert T == schedule["statistics"]["makespan"]

    def get_state(t, d):
        idx = 0
        while idx < len(d) and d[idx]["t"] < t:
            idx += 1
        if idx == 0:
            return np.array([float(d[0]["x"]), float(d[0]["y"])])
        if idx < len(d):
            last = np.array([float(d[idx - 1]["x"]), float(d[idx - 1]["y"])])
            nxt = np.array([float(d[idx]["x"]), float(d[idx]["y"])])
            dt = d[idx]["t"] - d[idx - 1]["t"]
            return (nxt - last) * ((t - d[idx - 1]["t"]) / dt) + last
        return np.array([float(d[-1]["x"]), float(d[-1]["y"])])
    for i in range(int(T + 1) * 10):
        pos = np.array([get_state(i / 10, schedule["schedule"][n]) for n in names])
        assert (pos >= -0.5).all() and (pos[:, 0] <= inst.dimx - 0.5).all() and (pos[:, 1] <= inst.dimy - 0.5).all()
        dist = np.linalg.norm(pos[:, None, :] - pos[None, :, :], axis=2) + 10 * np.eye(len(names))
        assert dist.min() >= 0.7, "visualize.py would print COLLISION at frame %d" % i


def test_path_pool_gives_the_same_answers(capi, set8, set32, monkeypatch):
    """cbs / ecbs batches keep their paths in a device pool (mrp_pathpool_*: nodes are
    lists of row numbers, conflict tables are gathered on the device, replans write
    their path into a fresh row); MRP_HOST_POOL=0 is the driver that ships whole
    tables.  Same searches: status, cost, makespan, expansion counts and every
    path must be identical."""
    from libmultirobotplanning_b200 import solver
    cases = [(solver.CBS, set8[:96], dict(max_hl=300)),
             (solver.CBS, [i for i in set32 if i.n_agents <= 20][:40], dict(max_hl=300)),
             (solver.ECBS, [i for i in set32 if i.n_agents in (30, 50)][:24], dict(w=1.3, max_hl=500)),
             (solver.ECBS, set8[:64], dict(w=1.5, max_hl=300))]
    for algo, insts, kw in cases:
        monkeypatch.setenv("MRP_HOST_POOL", "0")
        ref = solver.solve_batch(algo, insts, **kw)
        monkeypatch.delenv("MRP_HOST_POOL")
        assert sum(r["status"] == 0 for r in ref) >= len(insts) // 3
        # default: replans cut into slices of 256 expansions per launch (instances move on
        # as soon as their own replans are done); 0: every launch runs to its end; 5: a
        # search is suspended to its state blob and resumed dozens of times
        for slice_ in (None, "0", "5"):
            if slice_ is not None:
                monkeypatch.setenv("MRP_HOST_SLICE", slice_)
            got = solver.solve_batch(algo, insts, **kw)
            monkeypatch.delenv("MRP_HOST_SLICE", raising=False)
            for a, b in zip(got, ref):
                for key in ("status", "cost", "makespan", "lower_bound", "hl_expanded", "ll_expanded"):
                    assert a[key] == b[key], (slice_, key)
                assert len(a["paths"]) == len(b["paths"])
                for pa, pb in zip(a["paths"], b["paths"]):
                    assert np.array_equal(np.asarray(pa), np.asarray(pb))


def test_sliced_searches_with_spill_on_100_agent_instances(capi, set32, monkeypatch):
    """100-agent ECBS instances have replans of thousands of expansions: OPEN and the node pool
    outgrow their shared-memory parts (768 entries / 1024 nodes) and live on in the state blob,
    and a search is suspended and resumed dozens of times at a slice of 64.  Same answers as
    with replans that run to their end and as with the round-1 driver (whole path tables per
    launch, general kernel excluded by nothing: it takes what the tile kernel hands back)."""
    from libmultirobotplanning_b200 import solver
    insts = [i for i in set32 if i.n_agents == 100][:6]
    kw = dict(w=1.3, max_hl=400)
    monkeypatch.setenv("MRP_HOST_POOL", "0")
    ref = solver.solve_batch(solver.ECBS, insts, **kw)
    monkeypatch.delenv("MRP_HOST_POOL")
    assert sum(r["status"] == 0 for r in ref) >= 4
    assert max(r["ll_expanded"] for r in ref) > 20000
    for slice_ in ("64", "0", "1000"):
        monkeypatch.setenv("MRP_HOST_SLICE", slice_)
        got = solver.solve_batch(solver.ECBS, insts, **kw)
        monkeypatch.delenv("MRP_HOST_SLICE")
        for a, b in zip(got, ref):
            for key in ("status", "cost", "makespan", "lower_bound", "hl_expanded", "ll_expanded"):
                assert a[key] == b[key], (slice_, key)
            for pa, pb in zip(a["paths"], b["paths"]):
                assert np.array_equal(np.asarray(pa), np.asarray(pb))


def test_total_expansion_budget_caps_an_instance(capi, set32):
    """max_ll_total: an instance whose replans have used more low-level expansions than the
    budget is given up as capped (status 2) at its next high-level expansion; a generous budget
    changes nothing."""
    from libmultirobotplanning_b200 import solver
    insts = [i for i in set32 if i.n_agents == 50][:4]
    ref = solver.solve_batch(solver.ECBS, insts, w=1.3, max_hl=500)
    assert all(r["status"] == 0 and r["hl_expanded"] > 1 for r in ref)
    tight = solver.solve_batch(solver.ECBS, insts, w=1.3, max_hl=500, max_ll_total=100)
    assert all(r["status"] == 2 for r in tight)
    wide = solver.solve_batch(solver.ECBS, insts, w=1.3, max_hl=500, max_ll_total=10 ** 8)
    assert [(r["status"], r["cost"], r["ll_expanded"]) for r in wide] == \
        [(r["status"], r["cost"], r["ll_expanded"]) for r in ref]


def _ta_reference_cases(set8, set32, prefix):
    """(instance, answer of the unmodified reference cbs_ta binary) for the golden entries whose
    key starts with `prefix` (tests/golden/make_ref_golden_ta.py)."""
    import json
    import sys
    gdir = os.path.join(ROOT, "tests", "golden")
    if gdir not in sys.path:
        sys.path.insert(0, gdir)
    import make_ref_golden_ta as T
    g = json.load(open(os.path.join(gdir, "ref_binary_golden_ta.json")))["cbs_ta"]
    out = []
    for key, tool, inst, _, _ in T.selection(set8, set32):
        ref = g.get(key) if tool == "cbs_ta" and key.startswith(prefix) else None
        if ref is None or (ref["solved"] and ref["highLevelExpanded"] > 100):
            continue
        out.append((key, inst, ref))
    return out


def _solve_ta_by_shape(cases, max_hl):
    from libmultirobotplanning_b200 import solver
    res = [None] * len(cases)
    for dim in sorted({c[1].dimx for c in cases}):  # one batch per map size
        idx = [k for k, c in enumerate(cases) if c[1].dimx == dim]
        for k, r in zip(idx, solver.solve_batch(solver.CBS_TA, [cases[k][1] for k in idx],
                                                max_hl=max_hl)):
            res[k] = r
    return res


def test_cbs_ta_config_c4_equals_reference_binary(capi, set8, set32):
    """BASELINE config 4 (every goal of the file potential for every agent, 10-40 agents on
    32x32_obst204 and the 8x8 set): the optimal sum of costs equals what the UNMODIFIED
    reference cbs_ta prints (example/cbs_ta.cpp:589-600; golden made by
    tests/golden/make_ref_golden_ta.py from oracle/_ref/cbs_ta)."""
    cases = _ta_reference_cases(set8, set32, "all/")
    assert len(cases) >= 300
    for (key, inst, ref), r in zip(cases, _solve_ta_by_shape(cases, 20000)):
        assert ref["solved"] and r["status"] == 0 and r["cost"] == ref["cost"], (key, r["status"])
        if inst.n_agents <= 10:
            check_solution(inst, r["paths"], 1)


def test_cbs_ta_goal_subsets_equal_reference_binary(capi, set8, set32):
    """Seeded subsets of the goals per agent: agents with few or no potential goals (unassigned
    agents plan without a task, example/cbs_ta.cpp:283-319) and instances without any
    assignment / solution (the reference prints "Planning NOT successful!")."""
    cases = _ta_reference_cases(set8, set32, "sub/")
    assert len(cases) >= 100
    unsolved = 0
    for (key, inst, ref), r in zip(cases, _solve_ta_by_shape(cases, 20000)):
        if ref["solved"]:
            assert r["status"] == 0 and r["cost"] == ref["cost"], (key, r["status"])
        else:
            assert r["status"] == 1, (key, r["status"])
            unsolved += 1
    assert unsolved >= 1
