"""CPU tests of the host layer: the YAML reader of the drop-in binaries, their
flags / exit codes (example/cbs.cpp:571-596), and the loud failure without GPU."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "bin")


def test_yaml_reader_both_styles(tmp_path, set8, set32):
    from libmultirobotplanning_b200 import instances as I, solver
    for inst in (set8[3], set8[1999], set32[0]):
        p = str(tmp_path / "a.yaml")
        I.save_yaml(inst, p)  # 4-space block style with "-   goal:" items
        d = solver.load_instance_cli_parser(p)
        ref = I.load_yaml(p)   # PyYAML as the independent reader
        assert (d["dimx"], d["dimy"]) == (ref.dimx, ref.dimy)
        assert np.array_equal(d["obstacles"], ref.obstacles)
        assert np.array_equal(d["starts"], ref.cell(ref.starts))
        assert np.array_equal(d["goals"], ref.cell(ref.goals))
    # 2-space style of the reference's test fixtures, flow sequences, comments
    p = str(tmp_path / "b.yaml")
    open(p, "w").write("map:\n  dimensions: [5, 2]\n  obstacles:\n    - [0, 1]\n    - [1, 1]\n"
                       "agents:\n  - name: agent0\n    start: [0, 0]\n    goal: [4, 0]\n"
                       "  - name: agent1 # c\n    start: [1, 0]\n    goal: [3, 0]\n")
    d = solver.load_instance_cli_parser(p)
    assert d["starts"].tolist() == [0, 1] and d["goals"].tolist() == [4, 3]
    open(p, "w").write("map: {}\n")
    with pytest.raises(Exception):
        solver.load_instance_cli_parser(p)
    p = str(tmp_path / "c.yaml")
    open(p, "w").write("agents:\n- name: a\n  start: [0, 0]\n  potentialGoals: [[1, 0], [1, 0], [2, 0]]\n"
                       "- name: b\n  start: [1, 0]\n  potentialGoals: []\n"
                       "map:\n  dimensions: [3, 1]\n  obstacles: []\n")
    d = solver.load_instance_cli_parser(p, ta=True)
    assert [g.tolist() for g in d["potential_goals"]] == [[1, 2], []]


@pytest.mark.parametrize("tool", ["cbs", "ecbs", "cbs_ta", "ecbs_ta"])
def test_cli_flags(tool):
    exe = os.path.join(BIN, tool)
    r = subprocess.run([exe, "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--input" in r.stdout and "--output" in r.stdout
    assert ("--suboptimality" in r.stdout) == (tool in ("ecbs", "ecbs_ta"))
    assert ("--maxTaskAssignments" in r.stdout) == (tool in ("cbs_ta", "ecbs_ta"))
    r = subprocess.run([exe, "-i", "x.yaml"], capture_output=True, text=True)
    assert r.returncode == 1 and "required but missing" in r.stderr
    r = subprocess.run([exe, "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1 and "unrecognised option" in r.stderr


def test_cli_fails_loudly_without_gpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = str(tmp_path / "in.yaml")
    open(p, "w").write("map:\n  dimensions: [2, 1]\n  obstacles: []\nagents:\n"
                       "  - name: a\n    start: [0, 0]\n    goal: [1, 0]\n")
    out = str(tmp_path / "out.yaml")
    r = subprocess.run([os.path.join(BIN, "cbs"), "-i", p, "-o", out], capture_output=True,
                       text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    assert not os.path.exists(out)


def _write_movingai(tmp_path):
    rows = ["....@...", ".T......", "........", "..O....G", "S......W", "........"]
    (tmp_path / "t.map").write_text("type octile\nheight 6\nwidth 8\nmap\n" + "\n".join(rows) + "\n")
    scen = ["version 1"]
    pairs = [(1, (0, 0), (7, 5)), (0, (7, 0), (0, 5)), (0, (3, 2), (5, 2)), (2, (0, 2), (7, 2))]
    for b, s, g in pairs:
        scen.append("\t".join(map(str, [b, "t.map", 8, 6, s[0], s[1], g[0], g[1], 7.0])))
    (tmp_path / "t.scen").write_text("\n".join(scen) + "\n")
    return str(tmp_path / "t.scen"), str(tmp_path / "t.map")


def test_movingai_ingestion(tmp_path):
    """movingai .map/.scen reader: slicing and naming of the reference's
    standard_benchmark_converter.py, obstacles '@', 'T', 'O' kept."""
    from libmultirobotplanning_b200 import instances as I
    scen, mp = _write_movingai(tmp_path)
    w, h, obst = I.load_movingai_map(mp)
    assert (w, h) == (8, 6)
    assert sorted(map(tuple, obst.tolist())) == [(1, 1), (2, 3), (4, 0)]
    pairs = I.load_movingai_scen(scen, w, h, obst)
    # sorted by bucket, stable inside a bucket
    assert pairs == [((7, 0), (0, 5)), ((3, 2), (5, 2)), ((0, 0), (7, 5)), ((0, 2), (7, 2))]
    insts = I.movingai_instances(scen, mp, min_agents=2, agent_step=1)
    assert [i.n_agents for i in insts] == [2, 3, 4]
    assert insts[0].name.endswith("_2_agents")
    # round trip through the YAML the CLIs read
    out = str(tmp_path / "x.yaml")
    I.save_yaml(insts[2], out)
    back = I.load_yaml(out)
    assert np.array_equal(back.starts, insts[2].starts) and np.array_equal(back.goals, insts[2].goals)
    assert np.array_equal(back.obstacles, insts[2].obstacles) and (back.dimx, back.dimy) == (8, 6)
    # malformed inputs
    (tmp_path / "bad.map").write_text("type octile\nheight 1\nwidth 2\nmap\n.x\n")
    with pytest.raises(ValueError):
        I.load_movingai_map(str(tmp_path / "bad.map"))
    (tmp_path / "bad.scen").write_text("version 1\n0\tt.map\t8\t6\t4\t0\t1\t1\t1.0\n")
    with pytest.raises(ValueError):
        I.load_movingai_scen(str(tmp_path / "bad.scen"), 8, 6, obst)  # start on '@'


def test_movingai_oracle_solves(tmp_path, orc):
    from libmultirobotplanning_b200 import instances as I
    scen, mp = _write_movingai(tmp_path)
    inst = I.movingai_instances(scen, mp, min_agents=4, agent_step=1)[0]
    r = orc.cbs(inst.dimx, inst.dimy, inst.obstacles, inst.starts, inst.goals)
    assert r["status"] == 0 and r["cost"] >= 12 + 12 + 2 + 7


def test_host_assignment_reference_vectors():
    """host/assignment.hpp (dense successive shortest paths, parallel Murty
    partition) against the known answers of the reference's own tests
    (test/test_next_best_assignment.py:19-110)."""
    from libmultirobotplanning_b200 import solver
    nba = solver.next_best_assignments
    c, s = nba([], 0, 0)
    assert len(c) == 0
    c, s = nba([[0, 0, 2], [0, 1, 1]], 1, 2)
    assert list(c) == [1, 2] and s[0][0] == 1 and s[1][0] == 0
    c, s = nba([[0, 0, 2], [1, 0, 1]], 2, 1)
    assert list(c) == [1, 2] and list(s[0]) == [-1, 0] and list(s[1]) == [0, -1]
    c, s = nba([[0, 0, 90], [0, 1, 76], [1, 0, 35], [1, 1, 85]], 2, 2)
    assert list(c) == [111, 175] and list(s[0]) == [1, 0] and list(s[1]) == [0, 1]
    M = [[90, 76, 75, 80], [35, 85, 55, 65], [125, 95, 90, 105], [45, 110, 95, 115]]
    E = [[i, j, M[i][j]] for i in range(4) for j in range(4)]
    c, s = nba(E, 4, 4)
    assert len(c) == 24 and c[0] == 275 and c[-1] == 400
    assert list(s[0]) == [3, 2, 1, 0] and list(s[-1]) == [2, 1, 0, 3]
    assert all(c[k] <= c[k + 1] for k in range(len(c) - 1))
    assert len({tuple(r) for r in s}) == 24


def test_host_assignment_matches_oracle(orc):
    """Same cost sequence as the oracle's restatement of assignment.hpp /
    next_best_assignment.hpp on random problems: square and rectangular, sparse
    (missing edges, agents without any edge), many ties; every enumerated
    solution is a valid matching of the announced cost, no solution twice."""
    from libmultirobotplanning_b200 import solver
    rng = np.random.default_rng(11)
    for trial in range(40):
        A, T = int(rng.integers(1, 9)), int(rng.integers(1, 9))
        density = rng.choice([0.35, 0.7, 1.0])
        hi = int(rng.choice([3, 20, 1000]))
        cost = {}
        for a in range(A):
            for t in range(T):
                if rng.random() < density:
                    cost[(a, t)] = int(rng.integers(0, hi))
        E = [[a, t, c] for (a, t), c in cost.items()]
        K = 60
        c1, s1 = solver.next_best_assignments(E, A, T, K)
        c2, s2 = orc.next_best_assignments(E, A, T, K)
        assert list(c1) == list(c2), (trial, E)
        seen = set()
        for c, s in zip(c1, s1):
            used = [t for t in s if t >= 0]
            assert len(used) == len(set(used))
            assert all((a, int(t)) in cost for a, t in enumerate(s) if t >= 0)
            assert sum(cost[(a, int(t))] for a, t in enumerate(s) if t >= 0) == c
            assert tuple(s) not in seen
            seen.add(tuple(s))
    # a C4-sized problem: the first solutions of a 60 x 60 matrix
    M = rng.integers(5, 60, (60, 60))
    E = [[a, t, int(M[a, t])] for a in range(60) for t in range(60)]
    c1, _ = solver.next_best_assignments(E, 60, 60, 3)
    c2, _ = orc.next_best_assignments(E, 60, 60, 3)
    assert list(c1) == list(c2)


def test_validate_paths_agrees_with_the_oracle_conflict_count():
    """validate.py (the vectorised solution checker bench.py runs over whole ECBS batches) against the
    oracle's conflict loops (example/cbs.cpp:335-386, cbs_ta.cpp:369-420) on random walks."""
    from libmultirobotplanning_b200 import validate
    from oracle import orc

    class I:
        dimx, dimy = 9, 7
        obstacles = np.zeros((0, 2), np.int32)
    rng = np.random.default_rng(5)
    seen = {True: 0, False: 0}
    for trial in range(300):
        n = int(rng.integers(2, 7))
        paths, lens = [], []
        for a in range(n):
            L = int(rng.integers(1, 9))
            x, y = int(rng.integers(0, 9)), int(rng.integers(0, 7))
            p = [(x, y, 0)]
            for t in range(1, L):
                dx, dy = [(0, 0), (1, 0), (-1, 0), (0, 1), (0, -1)][int(rng.integers(0, 5))]
                x, y = min(8, max(0, x + dx)), min(6, max(0, y + dy))
                p.append((x, y, t))
            paths.append(np.array(p))
            lens.append(L)
        I.starts = np.array([p[0][:2] for p in paths])
        T = max(lens)
        cell = np.zeros((n, T), np.int32)
        for a, p in enumerate(paths):
            c = p[:, 0] + 9 * p[:, 1]
            cell[a, :len(c)] = c
            cell[a, len(c):] = c[-1]
        for mode in (0, 1):
            want = orc.first_conflict(cell, np.array(lens, np.int32), 9, mode) is None
            got = validate.validate_paths(I, paths, mode, goals=None) is None
            assert got == want, (trial, mode, validate.validate_paths(I, paths, mode))
            seen[want] += 1
    assert seen[True] > 20 and seen[False] > 20
    # moves and obstacles
    I.starts = np.array([[0, 0]])
    assert "more than one cell" in validate.validate_paths(I, [np.array([(0, 0, 0), (2, 0, 1)])])
    I.obstacles = np.array([[1, 0]], np.int32)
    assert "obstacle" in validate.validate_paths(I, [np.array([(0, 0, 0), (1, 0, 1)])])
    assert "goal" in validate.validate_paths(I, [np.array([(0, 0, 0), (0, 1, 1)])], goals=[[3, 3]])


def test_host_assignment_equals_reference_next_best_assignment(orc):
    """tests/golden/nba_golden.json: the complete cost sequences that the reference's OWN
    NextBestAssignment enumerates (next_best_assignment.hpp:37-201 over assignment.hpp, compiled
    unmodified into oracle/_ref/next_best_assignment) on 48 seeded cost tables (square, rectangular,
    sparse, many ties; 637 solutions).  The host module (host/assignment.hpp) and the oracle's
    restatement must enumerate the same costs in the same order."""
    import json
    import os
    import sys
    from libmultirobotplanning_b200 import solver
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if gdir not in sys.path:
        sys.path.insert(0, gdir)
    import make_nba_golden as N
    g = json.load(open(os.path.join(gdir, "nba_golden.json")))
    probs = N.problems()
    assert len(probs) == len(g) == 48
    for (A, T, E), want in zip(probs, g):
        edges = [list(e) for e in E]
        c1, _ = solver.next_best_assignments(edges, A, T, 500)
        c2, _ = orc.next_best_assignments(edges, A, T, 500)
        assert list(c1) == want and list(c2) == want, (A, T, E)
