"""Distance fields from the reference's OWN ShortestPathHeuristic class
(example/shortest_path_heuristic.hpp, compiled unmodified into oracle/_ref/sph_fields by
oracle/ref_build/Makefile): CRC-32 of the int32 fields per case -> tests/golden/sph_fields_golden.json.

    make -C oracle/ref_build && python tests/golden/make_sph_golden.py

Cases: benchmark maps with the goals of their agents, and seeded random maps of odd shapes
(1-wide, non-square, goals on obstacles, walled-in pockets).  The oracle's per-goal BFS and the
CUDA fields are compared with these sums in tests/.  Only runs where /root/reference exists."""
import json
import os
import subprocess
import sys
import tempfile
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from libmultirobotplanning_b200 import instances as I  # noqa: E402

EXE = os.path.join(ROOT, "oracle", "_ref", "sph_fields")
ODD_SHAPES = [(1, 9), (9, 1), (2, 2), (13, 7), (7, 13), (31, 33), (33, 31), (40, 33), (17, 45), (64, 5)]


def cases(s8, s32):
    """name -> (dimx, dimy, obstacles [n, 2], goals [g, 2]); deterministic"""
    out = {}
    for inst in s8[::100] + s32[::50]:
        out[inst.name] = (inst.dimx, inst.dimy, np.asarray(inst.obstacles, np.int32).reshape(-1, 2),
                          np.asarray(inst.goals, np.int32).reshape(-1, 2))
    rng = np.random.default_rng(20260101)
    for k, (dx, dy) in enumerate(ODD_SHAPES):
        for dens in (0.0, 0.2, 0.4):
            blocked = rng.random((dy, dx)) < dens
            ys, xs = np.nonzero(blocked)
            cells = rng.choice(dx * dy, min(dx * dy, 6), replace=False)  # may hit obstacles
            goals = np.stack([cells % dx, cells // dx], 1)
            out["odd_%dx%d_d%02d" % (dx, dy, int(dens * 100))] = (
                dx, dy, np.stack([xs, ys], 1).astype(np.int32), goals.astype(np.int32))
    return out


def reference_fields(dimx, dimy, obst, goals):
    txt = "%d %d %d %d\n" % (dimx, dimy, len(obst), len(goals))
    txt += "".join("%d %d\n" % (x, y) for x, y in obst) + "".join("%d %d\n" % (x, y) for x, y in goals)
    with tempfile.TemporaryDirectory() as td:  # the class drops searchGraph.dot into its cwd
        raw = subprocess.run([EXE], input=txt.encode(), stdout=subprocess.PIPE, check=True, cwd=td).stdout
    return np.frombuffer(raw, np.int32).reshape(len(goals), dimy * dimx)


def crc(fields):
    return zlib.crc32(np.ascontiguousarray(fields, np.int32).tobytes()) & 0xFFFFFFFF


if __name__ == "__main__":
    s8 = I.load_set(os.path.join(HERE, "bench_8x8.npz"))
    s32 = I.load_set(os.path.join(HERE, "bench_32x32.npz"))
    g = {}
    for name, (dx, dy, obst, goals) in cases(s8, s32).items():
        f = reference_fields(dx, dy, obst, goals)
        g[name] = {"crc32": crc(f), "goals": len(goals), "cells": dx * dy,
                   "unreachable": int((f == 2147483647).sum())}
    with open(os.path.join(HERE, "sph_fields_golden.json"), "w") as fo:
        json.dump(g, fo, indent=0, sort_keys=True)
    print(len(g), "cases")
