"""Runs the UNMODIFIED reference cbs_ta / ecbs_ta (example/cbs_ta.cpp, example/ecbs_ta.cpp with the
reference's own include/, built by oracle/ref_build/Makefile against stand-in Boost / yaml-cpp
headers into oracle/_ref/) over task-assignment instances and records their answers:
tests/golden/ref_binary_golden_ta.json.

    make -C oracle/ref_build && python tests/golden/make_ref_golden_ta.py

Instances: the benchmark files with `potentialGoals` = every goal of the file for every agent
(BASELINE config 4), and the same files with a seeded subset of the goals per agent (agents
with few or no potential goals: the maxTaskAssignments / unassigned-agent paths of
example/cbs_ta.cpp:283-319).  The optimal sum of costs of cbs_ta does not depend on how ties
are broken (in Boost.Heap or in the min-cost flow), so it is what the oracle and the CUDA path
are pinned to; ecbs_ta costs are recorded to be reported next to this path's.
Only runs where /root/reference exists (the build container)."""
import json
import os
import re
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
from libmultirobotplanning_b200 import instances as I  # noqa: E402
from make_ref_golden import run  # noqa: E402


def ex(inst):
    return int(re.search(r"ex(\d+)", inst.name).group(1))


def subset_goals(inst, seed, keep):
    """every agent keeps each goal of the file with probability `keep` (possibly none)"""
    rng = np.random.default_rng(seed)
    pg = [inst.goals[rng.random(len(inst.goals)) < keep] for _ in range(inst.n_agents)]
    return I.Instance(inst.name, inst.dimx, inst.dimy, inst.obstacles, inst.starts, None, pg)


def selection(s8, s32):
    """(key, tool, instance, extra args, timeout)"""
    jobs = []
    for inst in s32:  # config C4: every goal potential for every agent
        if inst.n_agents in (10, 20) or (inst.n_agents in (30, 40) and ex(inst) < 20):
            jobs.append(("all/" + inst.name, "cbs_ta", inst.with_all_goals_potential(), (), 60.0))
        if inst.n_agents in (10, 20) and ex(inst) < 10:
            jobs.append(("sub/" + inst.name, "cbs_ta", subset_goals(inst, 1000 + ex(inst), 0.3), (), 60.0))
            jobs.append(("all/" + inst.name, "ecbs_ta", inst.with_all_goals_potential(), ("-w", "1.3"), 60.0))
    for inst in s8:
        if ex(inst) < 10:
            jobs.append(("all/" + inst.name, "cbs_ta", inst.with_all_goals_potential(), (), 20.0))
            jobs.append(("sub/" + inst.name, "cbs_ta", subset_goals(inst, 2000 + ex(inst), 0.4), (), 20.0))
    return jobs


if __name__ == "__main__":
    s8 = I.load_set(os.path.join(HERE, "bench_8x8.npz"))
    s32 = I.load_set(os.path.join(HERE, "bench_32x32.npz"))
    jobs = selection(s8, s32)
    with ThreadPoolExecutor(max_workers=os.cpu_count()) as pool:
        res = list(pool.map(lambda j: run(j[1], j[2], j[3], timeout=j[4]), jobs))
    g = {"cbs_ta": {}, "ecbs_ta_w1.3": {}}
    timeouts = 0
    for (key, tool, _, _, _), r in zip(jobs, res):
        if r is None:
            timeouts += 1
            continue
        r.pop("runtime", None)
        g["cbs_ta" if tool == "cbs_ta" else "ecbs_ta_w1.3"][key] = r
    with open(os.path.join(HERE, "ref_binary_golden_ta.json"), "w") as f:
        json.dump(g, f, indent=0, sort_keys=True)
    print({k: len(v) for k, v in g.items()}, "timeouts:", timeouts)
