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
