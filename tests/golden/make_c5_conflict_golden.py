"""Golden answer of the config-C5 conflict sweep, from the CPU oracle alone.

The table is the one bench.py sweeps: 4096 agents on the synthetic 1024x1024 map,
agent i walks from its start down the distance field of goal i (steepest descent, neighbour preference
Left, Right, Up, Down), table int32 [4096][max_t + 1].  Fields: oracle BFS (oracle/mrp_oracle.cpp,
restating example/shortest_path_heuristic.hpp:12-62); conflicts: the oracle's restatement of
example/cbs.cpp:335-386 (first conflict) and example/ecbs.cpp:315-350 (count).
Writes tests/golden/c5_conflicts.json.  Takes a few minutes on the host cores.
usage: python tests/golden/make_c5_conflict_golden.py"""
import json, os, sys, time, zlib
from concurrent.futures import ThreadPoolExecutor
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from libmultirobotplanning_b200 import instances
from oracle import orc

DIM, N = 1024, 4096
inst = instances.synthetic_c5(dim=DIM, n_agents=N)
starts = (inst.starts[:N, 0] + DIM * inst.starts[:N, 1]).astype(np.int64)
INF = 2**31 - 1


def descend(field, start):
    """cells of the walk from start to the goal of `field` (Left, Right, Up, Down preference)"""
    cur = int(start)
    d = int(field[cur])
    path = [cur]
    while d > 0:
        x, y = cur % DIM, cur // DIM
        for dx, dy in ((-1, 0), (1, 0), (0, 1), (0, -1)):
            nx, ny = x + dx, y + dy
            if 0 <= nx < DIM and 0 <= ny < DIM and field[nx + DIM * ny] == d - 1:
                cur = nx + DIM * ny
                break
        else:
            raise RuntimeError("no descent")
        d -= 1
        path.append(cur)
    return path


def batch(b):
    g = inst.goals[b:b + 16]
    f = orc.bfs_fields(DIM, DIM, inst.obstacles, g)
    return [descend(f[k], starts[b + k]) for k in range(len(g))]


t0 = time.time()
with ThreadPoolExecutor(os.cpu_count() or 4) as ex:
    paths = [p for chunk in ex.map(batch, range(0, N, 16)) for p in chunk]
print("paths: %.1f s" % (time.time() - t0), flush=True)
length = np.array([len(p) for p in paths], np.int32)
T = int(length.max())
table = np.zeros((N, T), np.int32)
for a, p in enumerate(paths):
    table[a, :len(p)] = p
    table[a, len(p):] = p[-1]
t0 = time.time()
first = orc.first_conflict(table, length, DIM, 0)
count = orc.count_conflicts(table, length, 0)
print("oracle sweep: %.1f s" % (time.time() - t0), first, count, flush=True)
t, i, j, typ = first[0], first[1], first[2], first[3]  # (time, agent1, agent2, type, ...)
key = (t << 41) | (typ << 40) | (i << 20) | j
out = {"N": N, "T": T, "max_t": T - 1, "count": int(count), "first_conflict": [int(v) for v in first],
       "first_key": int(key), "table_crc32": zlib.crc32(table.tobytes()), "length_crc32": zlib.crc32(length.tobytes()),
       "how": "oracle BFS fields + steepest descent (Left, Right, Up, Down) + oracle all-pairs sweep, mode 0"}
with open(os.path.join(ROOT, "tests", "golden", "c5_conflicts.json"), "w") as f:
    json.dump(out, f, indent=1)
print(out)
