"""Conflict answers from the reference's OWN Environment methods (example/cbs.cpp:335-386,
example/cbs_ta.cpp:369-420, example/ecbs.cpp:282-350 and :401-452, included unmodified into the
oracle/_ref/env_probe_* binaries by oracle/ref_build/Makefile) on seeded random path tables ->
tests/golden/env_probe_golden.json.

    make -C oracle/ref_build && python tests/golden/make_env_golden.py

The tables are regenerated from their seeds by tests/ (tables() below); the file holds only the
reference's outputs: first conflict (time, agent1, agent2, type, x1, y1, x2, y2) or null under both
loop bounds, the conflict count of focalHeuristic, and focalState / focalTransition counts for
seeded candidate moves.  Only runs where /root/reference exists."""
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref")

# (dimx, dimy, agents, longest path): small maps so that pile-ups and swaps are frequent
SHAPES = [(6, 6, 8, 20), (3, 3, 5, 12), (12, 3, 10, 30), (32, 32, 60, 80), (32, 32, 100, 64),
          (2, 1, 2, 6), (5, 5, 2, 1), (8, 8, 30, 40), (40, 7, 17, 33), (4, 4, 12, 9)]
REPEATS = 4


def tables():
    """list of (dimx, cell[N][Tpad] int32, len[N] int32, queries[(self, t, from_cell, to_cell)])"""
    rng = np.random.default_rng(77)
    out = []
    moves = np.array([[0, 0], [1, 0], [-1, 0], [0, 1], [0, -1]])
    for dimx, dimy, N, T in SHAPES:
        for _ in range(REPEATS):
            ln = rng.integers(1, T + 1, N).astype(np.int32)
            ln[rng.integers(0, N)] = T
            cell = np.zeros((N, T), np.int32)
            for a in range(N):
                p = np.array([rng.integers(0, dimx), rng.integers(0, dimy)])
                for t in range(ln[a]):
                    cell[a, t] = p[0] + dimx * p[1]
                    p = np.clip(p + moves[rng.integers(0, 5)], 0, [dimx - 1, dimy - 1])
            qs = []
            for _ in range(6):
                s = int(rng.integers(0, N))
                t = int(rng.integers(0, T + 2))  # also past every path's end (getState clamps)
                o = int(rng.integers(0, N))      # start on another agent's cell: counts are non-trivial
                fc = int(cell[o, min(t, ln[o] - 1)])
                f = np.array([fc % dimx, fc // dimx])
                to = np.clip(f + moves[rng.integers(0, 5)], 0, [dimx - 1, dimy - 1])
                qs.append((s, t, fc, int(to[0] + dimx * to[1])))
            out.append((dimx, cell, ln, qs))
    return out


def probe(tool, tabs, with_queries):
    txt = ["%d" % len(tabs)]
    for dimx, cell, ln, qs in tabs:
        txt.append("%d" % len(ln))
        for a in range(len(ln)):
            c = cell[a, :ln[a]]
            txt.append("%d " % ln[a] + " ".join("%d %d" % (v % dimx, v // dimx) for v in c))
        if with_queries:
            txt.append("%d" % len(qs))
            for s, t, fc, tc in qs:
                txt.append("%d %d %d %d %d %d" % (s, t, fc % dimx, fc // dimx, tc % dimx, tc // dimx))
        else:
            txt.append("0")
    raw = subprocess.run([os.path.join(REF, tool)], input="\n".join(txt).encode(), stdout=subprocess.PIPE,
                         check=True, cwd="/tmp").stdout.decode().split("\n")
    res, cur = [], None
    for line in raw:
        w = line.split()
        if not w:
            continue
        if w[0] == "F":
            cur = {"first": [int(v) for v in w[2:]] if w[1] == "1" else None, "q": []}
            res.append(cur)
        elif w[0] == "C":
            cur["count"] = int(w[1])
        elif w[0] == "Q":
            cur["q"].append([int(w[1]), int(w[2])])
    return res


if __name__ == "__main__":
    tabs = tables()
    cbs = probe("env_probe_cbs", tabs, False)
    ecbs = probe("env_probe_ecbs", tabs, True)
    ta = probe("env_probe_cbs_ta", tabs, False)
    assert all(a["first"] == b["first"] for a, b in zip(cbs, ecbs))  # the two files hold the same loop
    g = {"first_mode0": [r["first"] for r in cbs], "first_mode1": [r["first"] for r in ta],
         "count": [r["count"] for r in ecbs], "focal_queries": [r["q"] for r in ecbs]}
    with open(os.path.join(HERE, "env_probe_golden.json"), "w") as f:
        json.dump(g, f)
    print(len(tabs), "tables;", sum(r is not None for r in g["first_mode0"]), "with a conflict (mode 0);",
          sum(r is not None for r in g["first_mode1"]), "(mode 1); counts up to", max(g["count"]))
