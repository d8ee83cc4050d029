"""CPU tests of the batched host drivers (libmultirobotplanning_b200/host/hl_search.hpp: the CBS /
ECBS / CBS-TA / ECBS-TA loops, flights, sliced replans, pool rows, lanes) on a machine without a
GPU.  The driver code under test is the product's libmrp_host.so, unchanged; what it calls is a
host emulation of the C ABI built from tests/emu/mrp_emu.cpp on top of the CPU oracle, loaded only
inside the subprocess started here (tests/emu/run_driver.py).  The answers are compared with the
committed answers of the unmodified reference binaries (tests/golden/ref_binary_golden*.json) and
with the reference's own fixtures; the switches that must not change a single answer (slices,
pool, lanes) are flipped against each other.  The same comparisons run against the CUDA library
in tests/test_gpu_solvers.py."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU_SRC = os.path.join(ROOT, "tests", "emu", "mrp_emu.cpp")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "_build", "libmrp_b200.so")
HOST_LIB = os.path.join(ROOT, "libmultirobotplanning_b200", "libmrp_host.so")
CBS, ECBS, CBS_TA, ECBS_TA = 0, 1, 2, 3


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(HOST_LIB):
        pytest.fail("libmrp_host.so is missing: run __graft_entry__.build()")
    srcs = [EMU_SRC, os.path.join(ROOT, "oracle", "mrp_oracle.cpp"), os.path.join(ROOT, "oracle", "mrp_oracle.h"),
            os.path.join(ROOT, "include", "mrp_b200.h")]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(s) > os.path.getmtime(EMU_LIB) for s in srcs):
        os.makedirs(os.path.dirname(EMU_LIB), exist_ok=True)
        cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
        subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-Wall", EMU_SRC,
                        os.path.join(ROOT, "oracle", "mrp_oracle.cpp"), "-o", EMU_LIB,
                        "-Wl,-soname,libmrp_b200.so"], check=True)

    def run(runs):
        p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "emu", "run_driver.py")],
                           input=json.dumps({"emu": EMU_LIB, "runs": runs}).encode(),
                           stdout=subprocess.PIPE, timeout=900)
        assert p.returncode == 0
        return json.loads(p.stdout)
    return run


def _agents(name):
    return int(re.search(r"agents(\d+)_", name).group(1))


def _check_paths(inst, paths, mode):
    from libmultirobotplanning_b200 import validate
    assert validate.validate_paths(inst, [np.array(p) for p in paths], mode) is None


def test_cbs_costs_and_switches(emu, set8, set32):
    """CBS through the driver: optimal sums of costs equal the unmodified reference binary's
    (cbs.hpp:85-172); slices on / tiny / off, pool off and one lane give the same answers."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_binary_golden.json")))["cbs"]
    n8 = [n for n in sorted(g) if "8by8" in n and _agents(n) >= 4 and g[n]["highLevelExpanded"] <= 400][:120]
    n32 = [n for n in sorted(g) if "32by32" in n and g[n]["highLevelExpanded"] <= 60][:12]
    base = {"algo": CBS, "max_hl": 20000}
    runs = [dict(base, set="bench_8x8", names=n8),
            dict(base, set="bench_8x8", names=n8, env={"MRP_HOST_SLICE": "5", "MRP_HOST_LANES": "1"}),
            dict(base, set="bench_8x8", names=n8, env={"MRP_HOST_SLICE": "0"}),
            dict(base, set="bench_8x8", names=n8, env={"MRP_HOST_POOL": "0", "MRP_HOST_LANES": "3"}),
            dict(base, set="bench_32x32", names=n32)]
    out = emu(runs)
    ref = [r for r in out[0]["results"]]
    for names, o in ((n8, out[0]), (n8, out[1]), (n8, out[2]), (n8, out[3]), (n32, out[4])):
        for n, r in zip(names, o["results"]):
            assert r["status"] == 0 and r["cost"] == g[n]["cost"], n
    for o in out[1:4]:  # the switches change nothing: same trees, node for node
        for a, b in zip(ref, o["results"]):
            assert (a["cost"], a["makespan"], a["hl_expanded"], a["ll_expanded"]) == \
                   (b["cost"], b["makespan"], b["hl_expanded"], b["ll_expanded"])
            assert a["paths"] == b["paths"]
    # what the driver used: default = pool rows + sliced launches; tiny slices suspend and resume
    assert out[0]["counters"][4] > 0 and out[0]["counters"][0] == 0
    assert out[1]["counters"][5] > 0 and out[1]["counters"][6] == out[1]["counters"][5]
    assert out[2]["counters"][4] == 0 and out[2]["counters"][3] > 0
    assert out[3]["counters"][2] > 0 and out[3]["counters"][1] == 0
    by = {i.name: i for i in set8 + set32}
    for n, r in list(zip(n8, out[1]["results"]))[::10] + list(zip(n32, out[4]["results"]))[::4]:
        _check_paths(by[n], r["paths"], 0)


def test_ecbs_bound_and_switches(emu, set32):
    """ECBS w = 1.3 (ecbs.hpp:109-288): solved, valid, cost <= w * lower bound, optimum <= cost;
    sliced roots / flights against the lock-step driver: identical answers."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_binary_golden.json")))
    names = sorted(g["ecbs_w1.3"])
    base = {"algo": ECBS, "w": 1.3, "max_hl": 2000, "set": "bench_32x32", "names": names}
    out = emu([base, dict(base, env={"MRP_HOST_SLICE": "16", "MRP_HOST_LANES": "2"}),
               dict(base, env={"MRP_HOST_SLICE": "0"}), dict(base, env={"MRP_HOST_POOL": "0"})])
    by = {i.name: i for i in set32}
    for n, r in zip(names, out[0]["results"]):
        assert r["status"] == 0, n
        assert np.float32(r["cost"]) <= np.float32(r["lower_bound"]) * np.float32(1.3), n
        if n in g["cbs"]:
            assert g["cbs"][n]["cost"] <= r["cost"], n
        assert r["cost"] <= 1.3 * g["ecbs_w1.3"][n]["cost"] + 1e-6, n  # both within w of one optimum
        _check_paths(by[n], r["paths"], 0)
    for o in out[1:]:
        for a, b in zip(out[0]["results"], o["results"]):
            assert (a["cost"], a["lower_bound"], a["hl_expanded"], a["ll_expanded"], a["paths"]) == \
                   (b["cost"], b["lower_bound"], b["hl_expanded"], b["ll_expanded"], b["paths"])
    assert out[1]["counters"][5] > 0


def test_cbs_ta_and_ecbs_ta(emu, set8, set32):
    """CBS-TA (cbs_ta.hpp:87-214) on config-C4 files and goal subsets: the optimal sums of costs and
    the "no solution" answers of the unmodified reference cbs_ta; ECBS-TA at w = 1.0 on the
    reference's fixtures (test/test_ecbs_ta.py:25-39) and within w on C4 files."""
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_binary_golden_ta.json")))["cbs_ta"]
    keys = [k for k in sorted(g) if not g[k]["solved"] or g[k]["highLevelExpanded"] <= 40]
    keys = [k for k in keys if "8by8" in k][::4] + [k for k in keys if "32by32" in k][::12]
    keys8 = [k for k in keys if "8by8" in k]
    keys32 = [k for k in keys if "32by32" in k]
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_fixtures.json")))
    fnames = [n for n in sorted(fx) if "cbs_ta_cost" in fx[n]["expected"]]
    out = emu([{"algo": CBS_TA, "ta": keys8, "max_hl": 20000}, {"algo": CBS_TA, "ta": keys32, "max_hl": 20000},
               {"algo": ECBS_TA, "fixtures": fnames, "w": 1.0, "max_hl": 5000},
               {"algo": CBS_TA, "fixtures": fnames, "max_hl": 5000},
               {"algo": ECBS_TA, "ta": [k for k in keys32 if g[k]["solved"]][:6], "w": 1.3, "max_hl": 5000},
               # at w = 1.0 ecbs_ta must return the optimum the reference cbs_ta found (test_ecbs_ta.py:25-39
               # asserts exactly that on the three fixtures)
               {"algo": ECBS_TA, "ta": [k for k in keys8 if g[k]["solved"]][:25], "w": 1.0, "max_hl": 20000}])
    unsolved = 0
    for ks, o in ((keys8, out[0]), (keys32, out[1])):
        # the runner returns the instances in the generator's order, not in the order of `ks`
        order = [k for k in _ta_order(set8, set32) if k in set(ks)]
        for k, r in zip(order, o["results"]):
            if g[k]["solved"]:
                assert r["status"] == 0 and r["cost"] == g[k]["cost"], k
            else:
                assert r["status"] == 1, k
                unsolved += 1
    assert unsolved >= 1 and len(keys8) >= 40 and len(keys32) >= 15
    for o in out[2:4]:
        for n, r in zip(fnames, o["results"]):
            assert r["status"] == 0 and r["cost"] == fx[n]["expected"]["cbs_ta_cost"], n
    order = [k for k in _ta_order(set8, set32) if k in set([k for k in keys32 if g[k]["solved"]][:6])]
    for k, r in zip(order, out[4]["results"]):
        assert r["status"] == 0 and g[k]["cost"] <= r["cost"] <= 1.3 * g[k]["cost"] + 1e-6, k
    order = [k for k in _ta_order(set8, set32) if k in set([k for k in keys8 if g[k]["solved"]][:25])]
    for k, r in zip(order, out[5]["results"]):
        assert r["status"] == 0 and r["cost"] == g[k]["cost"], k


def _ta_order(set8, set32):
    """keys of the cbs_ta golden in the order the generator (and the runner) lists the instances"""
    gdir = os.path.join(ROOT, "tests", "golden")
    if gdir not in sys.path:
        sys.path.insert(0, gdir)
    import make_ref_golden_ta as T
    return [j[0] for j in T.selection(set8, set32) if j[1] == "cbs_ta"]


def _emu_env():
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = os.path.dirname(EMU_LIB) + os.pathsep + env.get("LD_LIBRARY_PATH", "")
    return env


def test_cli_binaries_on_the_emulation(emu, ref_fixtures, set8, tmp_path):
    """bin/cbs, ecbs, cbs_ta, ecbs_ta (the drop-in command lines) with the emulation in front of
    their library path: the reference's pinned answers, the layout of output.yaml
    (example/cbs.cpp:637-661), "Planning NOT successful!" without an output file, option errors
    -> usage on stderr and exit status 1 (example/cbs.cpp:585-596); bin/mapf_sweep over 40 files."""
    import yaml
    from libmultirobotplanning_b200 import instances as I
    BIN = os.path.join(ROOT, "bin")
    env = _emu_env()
    for name, tool, extra, key in (("mapf_simple1", "cbs", [], "cbs_cost"),
                                   ("mapf_circle", "cbs", [], "cbs_cost"),
                                   ("mapf_atGoal", "ecbs", ["-w", "1.0"], "ecbs_w1_cost"),
                                   ("mapf_simple1", "ecbs", ["--suboptimality=1.0"], "ecbs_w1_cost"),
                                   ("mapfta_simple1_a2", "cbs_ta", [], "cbs_ta_cost"),
                                   ("mapfta_simple1_a3", "cbs_ta", ["--maxTaskAssignments", "5"], "cbs_ta_cost"),
                                   ("mapfta_simple1_a2", "ecbs_ta", ["-w", "1.0"], "cbs_ta_cost"),
                                   ("mapfta_simple1_a1", "ecbs_ta", [], "cbs_ta_cost")):
        d = ref_fixtures[name]
        inst = I.Instance(name, d["dimx"], d["dimy"], np.array(d["obstacles"], np.int32).reshape(-1, 2),
                          np.array(d["starts"], np.int32).reshape(-1, 2),
                          np.array(d["goals"], np.int32).reshape(-1, 2) if "goals" in d else None,
                          [np.array(p, np.int32).reshape(-1, 2) for p in d["potentialGoals"]]
                          if "potentialGoals" in d else None)
        inp, out = str(tmp_path / "in.yaml"), str(tmp_path / "out.yaml")
        if os.path.exists(out):
            os.remove(out)
        I.save_yaml(inst, inp)
        r = subprocess.run([os.path.join(BIN, tool), "-i", inp, "-o", out] + extra, capture_output=True,
                           text=True, env=env)
        assert r.returncode == 0, r.stderr
        assert "Planning successful!" in r.stdout and "done; cost:" in r.stdout
        text = open(out).read()
        y = yaml.safe_load(text)
        assert y["statistics"]["cost"] == d["expected"][key], (name, tool)
        assert list(y["statistics"].keys())[:5] == ["cost", "makespan", "runtime", "highLevelExpanded",
                                                    "lowLevelExpanded"]
        assert ("numTaskAssignments" in y["statistics"]) == (tool in ("cbs_ta", "ecbs_ta"))
        assert re.search(r"^schedule:\n  agent0:\n    - x: \d+\n      y: \d+\n      t: 0\n", text, re.M)
        if "agent0_last" in d["expected"]:
            assert y["schedule"]["agent0"][-1] == d["expected"]["agent0_last"]
    # an agent walled in: no solution, no output file, exit status 0 like the reference
    boxed = I.Instance("boxed", 3, 3, np.array([[1, 0], [0, 1], [1, 1]], np.int32), np.array([[0, 0]], np.int32),
                       np.array([[2, 2]], np.int32))
    inp, out = str(tmp_path / "boxed.yaml"), str(tmp_path / "boxed_out.yaml")
    I.save_yaml(boxed, inp)
    r = subprocess.run([os.path.join(BIN, "cbs"), "-i", inp, "-o", out], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "Planning NOT successful!" in r.stdout and not os.path.exists(out)
    r = subprocess.run([os.path.join(BIN, "cbs"), "-o", out], capture_output=True, text=True, env=env)
    assert r.returncode == 1 and "Allowed options" in r.stderr and "--input" in r.stderr
    r = subprocess.run([os.path.join(BIN, "ecbs"), "--help"], capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "suboptimality" in r.stdout
    # the streaming sweep: same costs as the reference binary's golden
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_binary_golden.json")))["cbs"]
    insts = [i for i in set8 if i.name in g and g[i.name]["highLevelExpanded"] <= 200][:40]
    files = []
    for i in insts:
        p = str(tmp_path / (i.name + ".yaml"))
        I.save_yaml(i, p)
        files.append(p)
    (tmp_path / "files.txt").write_text("\n".join(files) + "\n")
    (tmp_path / "out").mkdir()
    r = subprocess.run([os.path.join(BIN, "mapf_sweep"), "--algo", "cbs", "--batch", "16", "--maxHighLevelExpansions",
                        "5000", "--outputDir", str(tmp_path / "out"), "--csv", str(tmp_path / "res.csv"), "--list",
                        str(tmp_path / "files.txt")], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-500:]
    rows = [l.split(",") for l in (tmp_path / "res.csv").read_text().strip().splitlines()[1:]]
    assert [row[0] for row in rows] == files
    for row, i in zip(rows, insts):
        assert int(row[1]) == 0 and int(row[2]) == g[i.name]["cost"], i.name
        y = yaml.safe_load((tmp_path / "out" / (i.name + ".output.yaml")).read_text())
        assert y["statistics"]["cost"] == g[i.name]["cost"]


def test_reference_test_suite_on_our_binaries(emu, tmp_path):
    """Where /root/reference exists (the build container): the reference's OWN python tests
    (test/test_cbs.py, test_ecbs.py, test_cbs_ta.py, test_ecbs_ta.py: `./cbs -i ../test/... -o
    output.yaml`, 12 tests) run unmodified against THIS repository's command lines — the drop-in
    claim, checked by the reference's checker.  (PyYAML 6 needs an explicit Loader; the tests
    call yaml.load(f), so the launcher passes SafeLoader as its default.)"""
    ref = "/root/reference/test"
    if not os.path.isdir(ref):
        pytest.skip("the reference tree is not on this machine")
    build = tmp_path / "build"
    build.mkdir()
    os.symlink(ref, str(tmp_path / "test"))
    for b in ("cbs", "ecbs", "cbs_ta", "ecbs_ta"):
        os.symlink(os.path.join(ROOT, "bin", b), str(build / b))
    launcher = ("import yaml, functools, unittest\n"
                "yaml.load = functools.partial(yaml.load, Loader=yaml.SafeLoader)\n"
                "unittest.main(module=None, argv=['x', 'discover', '-s', '../test', '-p', 'test_*cbs*.py'])\n")
    r = subprocess.run([sys.executable, "-c", launcher], cwd=str(build), env=_emu_env(), capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert re.search(r"Ran 12 tests", r.stderr) and "OK" in r.stderr


def test_reference_templates_with_our_environment_on_the_emulation(emu, ref_fixtures, tmp_path):
    """oracle/_ref/templates_gpuenv — the reference's UNMODIFIED CBS / ECBS / CBSTA templates
    (include/libMultiRobotPlanning/cbs.hpp, ecbs.hpp, cbs_ta.hpp) instantiated with this
    repository's Environment adapter (host/gpu_environment.hpp) — with the emulation behind the
    adapter: the reference's own loops reach its pinned answers through our callbacks.  (The GPU
    run of the same binary is tests/test_gpu_solvers.py::test_reference_templates_drive_gpu_environment.)"""
    from libmultirobotplanning_b200 import instances as I
    exe = os.path.join(ROOT, "oracle", "_ref", "templates_gpuenv")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    n = 0
    for name, d in ref_fixtures.items():
        exp = d["expected"]
        inst = I.Instance(name, d["dimx"], d["dimy"], np.array(d["obstacles"], np.int32).reshape(-1, 2),
                          np.array(d["starts"], np.int32).reshape(-1, 2),
                          np.array(d["goals"], np.int32).reshape(-1, 2) if "goals" in d else None,
                          [np.array(p, np.int32).reshape(-1, 2) for p in d["potentialGoals"]]
                          if "potentialGoals" in d else None)
        inp = str(tmp_path / (name + ".yaml"))
        I.save_yaml(inst, inp)
        runs = []
        if "cbs_cost" in exp:
            runs = [("cbs", [], exp["cbs_cost"]), ("ecbs", ["1.0"], exp["ecbs_w1_cost"])]
        if "cbs_ta_cost" in exp:
            runs = [("cbs_ta", [], exp["cbs_ta_cost"])]
        for algo, extra, want in runs:
            r = subprocess.run([exe, algo, inp] + extra, capture_output=True, text=True, env=_emu_env())
            assert r.returncode == 0, (name, algo, r.stderr[-300:])
            assert ("cost %d" % want) in r.stdout, (name, algo, r.stdout[-200:])
            n += 1
    assert n >= 9


def test_caps_end_an_instance_as_capped(emu, set8, set32):
    """The reference's searches are unbounded; the drivers' caps (high-level expansions, expansions
    per replan, low-level expansions per instance, wall clock) must end an instance as CAPPED (2),
    never as "no solution", in the sliced, lock-step and table drivers — including flights that
    are dropped while their replans are suspended."""
    e100 = [i.name for i in set32 if i.n_agents == 100][:8]
    n8 = [i.name for i in set8 if i.n_agents >= 12][:40]
    out = emu([
        {"algo": ECBS, "set": "bench_32x32", "names": e100, "w": 1.3, "max_hl": 2000, "max_ll_total": 15000,
         "env": {"MRP_HOST_SLICE": "64", "MRP_HOST_LANES": "2", "MRP_HOST_LANE_SIZE": "4"}},
        {"algo": ECBS, "set": "bench_32x32", "names": e100, "w": 1.3, "max_hl": 3, "env": {"MRP_HOST_SLICE": "0"}},
        {"algo": CBS, "set": "bench_8x8", "names": n8, "max_hl": 40, "max_ll": 200},
        {"algo": CBS, "set": "bench_8x8", "names": n8, "max_hl": 40, "max_ll": 200,
         "env": {"MRP_HOST_SLICE": "0", "MRP_HOST_POOL": "0"}},
        {"algo": ECBS, "set": "bench_32x32", "names": e100, "w": 1.3, "max_hl": 2000, "max_seconds": 0.05,
         "env": {"MRP_HOST_SLICE": "64"}},
    ])
    for o in out:
        st = [r["status"] for r in o["results"]]
        assert set(st) <= {0, 2} and 2 in st
    # the same instances under the same caps: the sliced pool driver and the lock-step table driver agree
    assert [(r["status"], r["cost"]) for r in out[2]["results"]] == [(r["status"], r["cost"]) for r in out[3]["results"]]
