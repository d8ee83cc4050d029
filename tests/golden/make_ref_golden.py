"""Runs the UNMODIFIED reference cbs / ecbs (built against stand-in third-party
headers by oracle/ref_build/Makefile into oracle/_ref/) over benchmark instances
and records their answers: tests/golden/ref_binary_golden.json.

    make -C oracle/ref_build && python tests/golden/make_ref_golden.py

Only runs where /root/reference exists (the build container)."""
import json
import os
import re
import subprocess
import sys
import tempfile

import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from libmultirobotplanning_b200 import instances as I  # noqa: E402

REFBIN = os.path.join(ROOT, "oracle", "_ref")


def run(tool, inst, extra=(), timeout=10.0):
    with tempfile.TemporaryDirectory() as d:
        inp, out = os.path.join(d, "in.yaml"), os.path.join(d, "out.yaml")
        I.save_yaml(inst, inp)
        try:
            subprocess.run([os.path.join(REFBIN, tool), "-i", inp, "-o", out, *extra],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL,
                           timeout=timeout, check=True)
        except subprocess.TimeoutExpired:
            return None
        if not os.path.exists(out):
            return {"solved": False}
        with open(out) as f:
            y = yaml.safe_load(f)
        s = y["statistics"]
        r = {"solved": True, "cost": s["cost"], "makespan": s["makespan"],
             "highLevelExpanded": s["highLevelExpanded"],
             "lowLevelExpanded": s["lowLevelExpanded"], "runtime": s["runtime"]}
        if "numTaskAssignments" in s:  # cbs_ta / ecbs_ta (example/cbs_ta.cpp:599)
            r["numTaskAssignments"] = s["numTaskAssignments"]
        return r


if __name__ == "__main__":
    s8 = I.load_set(os.path.join(HERE, "bench_8x8.npz"))
    s32 = I.load_set(os.path.join(HERE, "bench_32x32.npz"))
    g = {"cbs": {}, "ecbs_w1.3": {}}
    ex = lambda i: int(re.search(r"ex(\d+)", i.name).group(1))
    for inst in s8:
        if inst.n_agents <= 5 or (inst.n_agents <= 9 and ex(inst) < 12):
            r = run("cbs", inst)
            if r:
                g["cbs"][inst.name] = r
    for inst in s32:
        if inst.n_agents == 10 and ex(inst) < 40:
            r = run("cbs", inst)
            if r:
                g["cbs"][inst.name] = r
        if inst.n_agents in (10, 30, 50) and ex(inst) < 12:
            r = run("ecbs", inst, ("-w", "1.3"), timeout=30.0)
            if r:
                g["ecbs_w1.3"][inst.name] = r
    for k in g:
        for r in g[k].values():
            r.pop("runtime", None)
    with open(os.path.join(HERE, "ref_binary_golden.json"), "w") as f:
        json.dump(g, f, indent=0, sort_keys=True)
    print({k: len(v) for k, v in g.items()})
