"""Regenerates tests/golden/* from the reference checkout (run in the build
container only; /root/reference does not exist on the GPU box).

    python tests/golden/make_golden.py [/root/reference]

Outputs
  bench_8x8.npz, bench_32x32.npz   the reference's benchmark/ instances, packed
  ref_fixtures.json                 the reference's test/*.yaml inputs together
                                    with the answers its own tests assert
  oracle_golden.json                answers of the CPU oracle on fixed cases
                                    (function-level golden vectors; not pinned
                                    by the reference, see oracle/mrp_oracle.h)
"""
import glob
import json
import os
import re
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from libmultirobotplanning_b200 import instances as I  # noqa: E402
from oracle import orc  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"


def natural(name):
    return [int(t) if t.isdigit() else t for t in re.split(r"(\d+)", name)]


def pack(dirname, out):
    files = sorted(glob.glob(os.path.join(REF, "benchmark", dirname, "*.yaml")),
                   key=lambda p: natural(os.path.basename(p)))
    insts = []
    for p in files:
        inst = I.load_yaml(p)
        inst.name = os.path.splitext(os.path.basename(p))[0]
        insts.append(inst)
    I.save_set(os.path.join(HERE, out), insts)
    return insts


def fixtures():
    # expected values: test/test_cbs.py:24-34, test/test_ecbs.py:25-35,
    # test/test_cbs_ta.py:24-38 (and test_ecbs_ta.py:25-39)
    expected = {
        "mapf_simple1": {"cbs_cost": 8, "ecbs_w1_cost": 8},
        "mapf_circle": {"cbs_cost": 4, "ecbs_w1_cost": 4},
        "mapf_atGoal": {"cbs_cost": 0, "ecbs_w1_cost": 0},
        "mapfta_simple1_a1": {"cbs_ta_cost": 6},
        "mapfta_simple1_a2": {"cbs_ta_cost": 6,
                              "agent0_last": {"x": 4, "y": 0, "t": 4},
                              "agent1_last_xy": [2, 1]},
        "mapfta_simple1_a3": {"cbs_ta_cost": 5,
                              "agent0_last": {"x": 3, "y": 0, "t": 3}},
        # not asserted for cbs by the reference (SIPP-only fixtures); kept as
        # extra inputs, answers come from the oracle
        "mapf_swap2": {}, "mapf_swap4": {}, "mapf_simple1b": {},
        "mapf_someAtGoal": {},
    }
    out = {}
    for name, exp in expected.items():
        inst = I.load_yaml(os.path.join(REF, "test", name + ".yaml"))
        d = {"dimx": inst.dimx, "dimy": inst.dimy,
             "obstacles": inst.obstacles.tolist(),
             "starts": inst.starts.tolist(), "expected": exp}
        if inst.goals is not None:
            d["goals"] = inst.goals.tolist()
        else:
            d["potentialGoals"] = [g.tolist() for g in inst.potential_goals]
        out[name] = d
    with open(os.path.join(HERE, "ref_fixtures.json"), "w") as f:
        json.dump(out, f, indent=1)
    return out


def crc(a):
    return zlib.crc32(np.ascontiguousarray(a, dtype="<i4").tobytes()) & 0xFFFFFFFF


def oracle_golden(s8, s32):
    g = {}
    # distance fields: CRC32 of int32[G][cells] for a spread of instances
    fields = {}
    for inst in s8[::97] + s32[::53]:
        f = orc.bfs_fields(inst.dimx, inst.dimy, inst.obstacles, inst.goals)
        fields[inst.name] = {"crc32": crc(f), "n_goals": int(len(inst.goals)),
                             "sum_finite": int(f[f != orc.INF].sum()),
                             "n_inf": int((f == orc.INF).sum())}
    g["bfs_fields"] = fields
    # CBS sums-of-costs: all 8x8 instances with <= 5 agents (cap independent,
    # all solved), a sample of harder ones under a cap, 32x32 10-agent sample
    caps = (20000, 2_000_000, 20.0)
    cbs = {}
    for inst in s8:
        n = inst.n_agents
        ex = int(re.search(r"ex(\d+)", inst.name).group(1))
        if n <= 5 or (n <= 8 and ex < 10):
            r = orc.cbs(inst.dimx, inst.dimy, inst.obstacles, inst.starts,
                        inst.goals, caps)
            cbs[inst.name] = {"status": r["status"], "cost": r["cost"],
                              "makespan": r["makespan"]}
    for inst in s32:
        if inst.n_agents == 10 and int(re.search(r"ex(\d+)", inst.name).group(1)) < 20:
            r = orc.cbs(inst.dimx, inst.dimy, inst.obstacles, inst.starts,
                        inst.goals, caps)
            cbs[inst.name] = {"status": r["status"], "cost": r["cost"],
                              "makespan": r["makespan"]}
    g["cbs"] = cbs
    with open(os.path.join(HERE, "oracle_golden.json"), "w") as f:
        json.dump(g, f, indent=0, sort_keys=True)
    return g


if __name__ == "__main__":
    orc.build()
    s8 = pack("8x8_obst12", "bench_8x8.npz")
    s32 = pack("32x32_obst204", "bench_32x32.npz")
    print("packed", len(s8), len(s32))
    fixtures()
    g = oracle_golden(s8, s32)
    by_n = {}
    for name, r in g["cbs"].items():
        m = re.match(r"map_8by8_obst12_agents(\d+)_ex", name)
        if m and r["status"] == 0:
            by_n.setdefault(int(m.group(1)), []).append(r["cost"])
    print({n: (len(v), sum(v)) for n, v in sorted(by_n.items())})
