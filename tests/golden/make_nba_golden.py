"""Cost sequences of the reference's OWN NextBestAssignment (include/libMultiRobotPlanning/
next_best_assignment.hpp + assignment.hpp; example/next_best_assignment.cpp compiled unmodified into
oracle/_ref/next_best_assignment against the Boost.Graph stand-in) on seeded random cost tables ->
tests/golden/nba_golden.json.

    make -C oracle/ref_build && python tests/golden/make_nba_golden.py

The SEQUENCE OF COSTS of the enumeration is unique (which of several equal-cost assignments comes first is
not); the oracle's restatement and the host module (host/assignment.hpp) must produce the same sequences.
Only runs where /root/reference exists."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
EXE = os.path.join(ROOT, "oracle", "_ref", "next_best_assignment")


def problems():
    """[(n_agents, n_tasks, edges [(agent, task, cost)])]: square, rectangular, sparse, many ties"""
    rng = np.random.default_rng(2718)
    out = []
    for _ in range(48):
        A, T = int(rng.integers(1, 6)), int(rng.integers(1, 6))
        density = float(rng.choice([0.4, 0.75, 1.0]))
        hi = int(rng.choice([2, 10, 500]))
        E = [(a, t, int(rng.integers(0, hi))) for a in range(A) for t in range(T) if rng.random() < density]
        out.append((A, T, E))
    return out


def reference_costs(E):
    with tempfile.TemporaryDirectory() as td:
        inp, outp = os.path.join(td, "in.txt"), os.path.join(td, "out.yaml")
        with open(inp, "w") as f:
            for a, t, c in E:
                f.write("a%d->t%d:%d\n" % (a, t, c))
        subprocess.run([EXE, "-i", inp, "-o", outp], check=True, stdout=subprocess.DEVNULL, timeout=300, cwd=td)
        y = yaml.safe_load(open(outp))
    return [s["cost"] for s in (y["solutions"] or [])]


if __name__ == "__main__":
    g = [reference_costs(E) for _, _, E in problems()]
    with open(os.path.join(HERE, "nba_golden.json"), "w") as f:
        json.dump(g, f)
    print(len(g), "problems; solutions per problem up to", max(len(x) for x in g))
