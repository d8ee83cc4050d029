"""MAPF instance containers, the input YAML reader and the synthetic generators.

Input format (example/cbs.cpp:598-618, example/cbs_ta.cpp:547-568 of the
reference): `map.dimensions: [dimx, dimy]`, `map.obstacles: [[x, y], ...]`,
`agents[]: {name, start: [x, y], goal: [x, y] | potentialGoals: [[x, y], ...]}`.

Packed instance sets (`tests/golden/bench_*.npz`) hold the reference's
`benchmark/` instances in a compact form so they can travel to the GPU box
(`/root/reference` does not exist there).
"""
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

MASK64 = (1 << 64) - 1


@dataclass
class Instance:
    name: str
    dimx: int
    dimy: int
    obstacles: np.ndarray            # [n_obst, 2] int32 (x, y)
    starts: np.ndarray               # [n_agents, 2] int32
    goals: Optional[np.ndarray] = None       # [n_agents, 2] int32
    potential_goals: Optional[List[np.ndarray]] = None  # cbs_ta

    @property
    def n_agents(self):
        return len(self.starts)

    def cell(self, xy):
        xy = np.asarray(xy)
        return xy[..., 0] + self.dimx * xy[..., 1]

    def with_all_goals_potential(self):
        """Config C4: every agent may take any of the instance's goals."""
        pg = [self.goals.copy() for _ in range(self.n_agents)]
        return Instance(self.name, self.dimx, self.dimy, self.obstacles,
                        self.starts, None, pg)


def load_yaml(path) -> Instance:
    import yaml
    with open(path) as f:
        cfg = yaml.safe_load(f)
    dimx, dimy = cfg["map"]["dimensions"]
    obst = np.array(cfg["map"].get("obstacles") or [], np.int32).reshape(-1, 2)
    starts, goals, pgs = [], [], []
    is_ta = False
    for a in cfg["agents"]:
        starts.append(a["start"])
        if "potentialGoals" in a:
            is_ta = True
            pgs.append(np.array(a["potentialGoals"] or [], np.int32).reshape(-1, 2))
        else:
            goals.append(a["goal"])
    starts = np.array(starts, np.int32).reshape(-1, 2)
    if is_ta:
        return Instance(str(path), dimx, dimy, obst, starts, None, pgs)
    return Instance(str(path), dimx, dimy, obst, starts,
                    np.array(goals, np.int32).reshape(-1, 2))


def save_yaml(inst: Instance, path):
    """Writes the 4-space block style the reference's benchmark files use."""
    with open(path, "w") as f:
        f.write("agents:\n")
        for i in range(inst.n_agents):
            if inst.goals is not None:
                f.write("-   goal: [%d, %d]\n" % tuple(inst.goals[i]))
                f.write("    name: agent%d\n" % i)
            else:
                f.write("-   name: agent%d\n" % i)
                pg = inst.potential_goals[i]
                if len(pg) == 0:
                    f.write("    potentialGoals: []\n")
                else:
                    f.write("    potentialGoals:\n")
                    for p in pg:
                        f.write("    - [%d, %d]\n" % tuple(p))
            f.write("    start: [%d, %d]\n" % tuple(inst.starts[i]))
        f.write("map:\n    dimensions: [%d, %d]\n" % (inst.dimx, inst.dimy))
        if len(inst.obstacles) == 0:
            f.write("    obstacles: []\n")
        else:
            f.write("    obstacles:\n")
            for o in inst.obstacles:
                f.write("    - [%d, %d]\n" % tuple(o))


def save_set(path, instances: List[Instance]):
    names = np.array([i.name for i in instances])
    dims = np.array([[i.dimx, i.dimy] for i in instances], np.int16)
    ooff = np.zeros(len(instances) + 1, np.int32)
    aoff = np.zeros(len(instances) + 1, np.int32)
    for k, i in enumerate(instances):
        ooff[k + 1] = ooff[k] + len(i.obstacles)
        aoff[k + 1] = aoff[k] + i.n_agents
    obst = np.concatenate([i.obstacles for i in instances]).astype(np.int16)
    starts = np.concatenate([i.starts for i in instances]).astype(np.int16)
    goals = np.concatenate([i.goals for i in instances]).astype(np.int16)
    np.savez_compressed(path, names=names, dims=dims, obst_off=ooff,
                        agent_off=aoff, obst=obst, starts=starts, goals=goals)


def load_set(path) -> List[Instance]:
    z = np.load(path, allow_pickle=False)
    out = []
    for k, name in enumerate(z["names"]):
        o0, o1 = z["obst_off"][k], z["obst_off"][k + 1]
        a0, a1 = z["agent_off"][k], z["agent_off"][k + 1]
        out.append(Instance(str(name), int(z["dims"][k, 0]), int(z["dims"][k, 1]),
                            z["obst"][o0:o1].astype(np.int32),
                            z["starts"][a0:a1].astype(np.int32),
                            z["goals"][a0:a1].astype(np.int32)))
    return out


# ---------------------------------------------------------------------------
# movingai.com MAPF benchmarks (.map / .scen) — SURVEY.md §8(f4).
# Same slicing and file naming as the reference's
# example/standard_benchmark_converter.py:30-89 (agents sorted by bucket, the
# first 10, 20, ... agents of a scenario per instance).  One deliberate
# difference: that script compares each map character with the whole SET of
# occupied characters (line 45), which is never true, so the YAML files it
# writes have no obstacles at all; here '@', 'T' and 'O' are obstacles, as the
# script's own `occupied_char` default intends.
# ---------------------------------------------------------------------------
MOVINGAI_OCCUPIED = frozenset("@TO")
MOVINGAI_VALID = frozenset("@.TGOSW")


def load_movingai_map(path):
    """-> (width, height, obstacles[n, 2] int32 as (x, y))."""
    with open(path) as f:
        lines = f.read().splitlines()
    hdr = {}
    k = 0
    while k < len(lines) and lines[k].strip() != "map":
        parts = lines[k].split()
        if len(parts) == 2:
            hdr[parts[0]] = parts[1]
        k += 1
    if "height" not in hdr or "width" not in hdr or k == len(lines):
        raise ValueError("%s: not a movingai .map file" % path)
    height, width = int(hdr["height"]), int(hdr["width"])
    rows = lines[k + 1:k + 1 + height]
    if len(rows) != height:
        raise ValueError("%s: %d map rows, header says %d" % (path, len(rows), height))
    obst = []
    for y, row in enumerate(rows):
        if len(row) != width:
            raise ValueError("%s: row %d has %d cells, header says %d" % (path, y, len(row), width))
        for x, c in enumerate(row):
            if c not in MOVINGAI_VALID:
                raise ValueError("%s: unknown map character %r at (%d, %d)" % (path, c, x, y))
            if c in MOVINGAI_OCCUPIED:
                obst.append((x, y))
    return width, height, np.array(obst, np.int32).reshape(-1, 2)


def load_movingai_scen(path, width, height, obstacles=None):
    """-> list of ((sx, sy), (gx, gy)), sorted by bucket like the reference's converter."""
    with open(path) as f:
        lines = [l for l in f.read().splitlines() if l.strip()]
    if not lines or "version 1" not in lines[0]:
        raise ValueError("%s: .scen version type does not match" % path)
    rows = []
    for l in lines[1:]:
        c = l.split("\t") if "\t" in l else l.split()
        if len(c) < 9:
            raise ValueError("%s: malformed scenario line %r" % (path, l))
        bucket, w, h = int(c[0]), int(c[2]), int(c[3])
        if (w, h) != (width, height):
            raise ValueError("%s: scenario is for a %dx%d map, not %dx%d" % (path, w, h, width, height))
        rows.append((bucket, (int(c[4]), int(c[5])), (int(c[6]), int(c[7]))))
    rows.sort(key=lambda r: r[0])  # stable: the order inside a bucket is kept
    blocked = set(map(tuple, np.asarray(obstacles).tolist())) if obstacles is not None else set()
    for _, s, g in rows:
        if s in blocked or g in blocked:
            raise ValueError("%s: start %s / goal %s lies on an obstacle" % (path, s, g))
    return [(s, g) for _, s, g in rows]


def movingai_instances(scen_path, map_path, min_agents=10, agent_step=10, max_agents=None):
    """One Instance per agent count min_agents, min_agents + agent_step, ... (the
    first k agents of the scenario), named like the converter's output files."""
    width, height, obst = load_movingai_map(map_path)
    pairs = load_movingai_scen(scen_path, width, height, obst)
    hi = len(pairs) if max_agents is None else min(len(pairs), max_agents)
    out = []
    for k in range(min_agents, hi + 1, agent_step):
        starts = np.array([p[0] for p in pairs[:k]], np.int32).reshape(-1, 2)
        goals = np.array([p[1] for p in pairs[:k]], np.int32).reshape(-1, 2)
        out.append(Instance("%s_%d_agents" % (scen_path, k), width, height, obst, starts, goals))
    return out


# ---------------------------------------------------------------------------
# Synthetic generators — stateless RNG so
# CPU and GPU sides can regenerate identical inputs.
# ---------------------------------------------------------------------------
def splitmix64(x):
    """Vectorised splitmix64 over uint64 numpy arrays (wrap-around)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _component_of(free, seed_xy):
    """Boolean mask of the 4-connected component of `free` containing seed."""
    dimy, dimx = free.shape
    comp = np.zeros_like(free)
    comp[seed_xy[1], seed_xy[0]] = free[seed_xy[1], seed_xy[0]]
    while True:
        grow = comp.copy()
        grow[1:, :] |= comp[:-1, :]
        grow[:-1, :] |= comp[1:, :]
        grow[:, 1:] |= comp[:, :-1]
        grow[:, :-1] |= comp[:, 1:]
        grow &= free
        if (grow == comp).all():
            return comp
        comp = grow


def synthetic_c5(dim=1024, density=0.20, n_agents=4096, seed=0xB2000005):
    """Config C5 map: `dim`x`dim`, obstacle iff
    (splitmix64(seed ^ (y*dim+x)) >> 11) * 2^-53 < density; starts then goals =
    first 2*n_agents distinct cells c_i = splitmix64(seed_agents + i) & (dim^2-1)
    inside the giant component (the one containing (dim/2+1, dim/2))."""
    assert dim & (dim - 1) == 0
    idx = np.arange(dim * dim, dtype=np.uint64)
    r = splitmix64(np.uint64(seed) ^ idx)
    u = (r >> np.uint64(11)).astype(np.float64) * (2.0 ** -53)
    blocked = (u < density).reshape(dim, dim)
    free = ~blocked
    seed_xy = (dim // 2 + 1, dim // 2)
    if not free[seed_xy[1], seed_xy[0]]:
        ys, xs = np.nonzero(free)
        k = np.argmin(np.abs(xs - seed_xy[0]) + np.abs(ys - seed_xy[1]))
        seed_xy = (int(xs[k]), int(ys[k]))
    comp = _component_of_fast(free, seed_xy)
    picked, seen = [], set()
    i = 0
    seed_a = 0xB2000055 if seed == 0xB2000005 else (seed ^ 0x50)
    while len(picked) < 2 * n_agents:
        batch = splitmix64(np.uint64(seed_a) + np.arange(i, i + 4096, dtype=np.uint64))
        i += 4096
        for c in (batch & np.uint64(dim * dim - 1)).astype(np.int64):
            c = int(c)
            if c in seen:
                continue
            seen.add(c)
            if comp[c // dim, c % dim]:
                picked.append(c)
                if len(picked) == 2 * n_agents:
                    break
    picked = np.array(picked, np.int64)
    xy = np.stack([picked % dim, picked // dim], 1).astype(np.int32)
    ys, xs = np.nonzero(blocked)
    obst = np.stack([xs, ys], 1).astype(np.int32)
    return Instance("synthetic_c5_%dx%d" % (dim, dim), dim, dim, obst,
                    xy[:n_agents], xy[n_agents:])


def _component_of_fast(free, seed_xy):
    """Queue flood fill (the vectorised fixed point above needs ~depth sweeps)."""
    dimy, dimx = free.shape
    comp = np.zeros(free.shape, bool)
    if not free[seed_xy[1], seed_xy[0]]:
        return comp
    f = free.ravel()
    cflat = comp.ravel()
    start = seed_xy[0] + dimx * seed_xy[1]
    cflat[start] = True
    frontier = np.array([start], np.int64)
    while len(frontier):
        x = frontier % dimx
        y = frontier // dimx
        cand = np.concatenate([frontier[x > 0] - 1, frontier[x < dimx - 1] + 1,
                               frontier[y > 0] - dimx,
                               frontier[y < dimy - 1] + dimx])
        cand = cand[f[cand] & ~cflat[cand]]
        cand = np.unique(cand)
        cflat[cand] = True
        frontier = cand
    return comp


def synthetic_c3(base: Instance, k: int, n_agents: int):
    """Config C3: keep the file's agents, append agents whose start/goal are
    drawn without replacement from the free cells of the goals' connected
    component with draw d = splitmix64(0xB2000003 ^ (k<<16) ^ d) mod (dimx*dimy)."""
    dimx, dimy = base.dimx, base.dimy
    free = np.ones((dimy, dimx), bool)
    free[base.obstacles[:, 1], base.obstacles[:, 0]] = False
    comp = _component_of_fast(free, tuple(base.goals[0]))
    used_s = set(map(int, base.cell(base.starts)))
    used_g = set(map(int, base.cell(base.goals)))
    starts = [tuple(s) for s in base.starts]
    goals = [tuple(g) for g in base.goals]
    d = 0
    want_start = True
    pending = None
    while len(goals) < n_agents:
        c = int(splitmix64(np.uint64(0xB2000003 ^ (k << 16) ^ d))) % (dimx * dimy)
        d += 1
        if not comp[c // dimx, c % dimx]:
            continue
        if want_start:
            if c in used_s:
                continue
            used_s.add(c)
            pending = (c % dimx, c // dimx)
            want_start = False
        else:
            if c in used_g:
                continue
            used_g.add(c)
            starts.append(pending)
            goals.append((c % dimx, c // dimx))
            want_start = True
    return Instance("%s_n%d" % (base.name, n_agents), dimx, dimy, base.obstacles,
                    np.array(starts[:n_agents], np.int32),
                    np.array(goals[:n_agents], np.int32))
