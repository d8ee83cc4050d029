"""ctypes binding of libmrp_host.so: the batched CBS / ECBS / CBS-TA drivers
(host/hl_search.hpp) that sit on top of the CUDA hot path.  Used by the tests
and bench.py; the user-facing entry points are the cbs / ecbs / cbs_ta binaries
under bin/."""
import ctypes as C
import os

import numpy as np

from . import _capi

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libmrp_host.so")
BIN_DIR = os.path.join(os.path.dirname(_PKG), "bin")
CBS, ECBS, CBS_TA, ECBS_TA = 0, 1, 2, 3
SOLVED, NO_SOLUTION, CAPPED = 0, 1, 2

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise _capi.MrpError(-1, "%s not found: run __graft_entry__.build()" % LIB_PATH)
        _capi.lib()  # libmrp_b200.so first (dependency)
        # the OpenMP team of the host driver sleeps between parallel regions:
        # spinning workers would fight the CUDA driver threads for the cores
        os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
        _lib = C.CDLL(LIB_PATH)
        _lib.mrph_last_error.restype = C.c_char_p
    return _lib


def next_best_assignments(edges, n_agents, n_tasks, max_solutions=1000):
    """The host assignment module on its own (CPU only): edges [n][3] =
    (agent, task, cost) -> (costs[k], assignment[k][agent] or -1) in
    non-decreasing cost, as NextBestAssignment::nextSolution of the reference."""
    e = np.ascontiguousarray(edges, dtype=np.int64).reshape(-1, 3)
    costs = np.zeros(max_solutions, np.int64)
    sol = np.full((max_solutions, max(n_agents, 1)), -1, np.int32)
    if not os.path.exists(LIB_PATH):
        raise _capi.MrpError(-1, "%s not found: run __graft_entry__.build()" % LIB_PATH)
    h = C.CDLL(LIB_PATH) if _lib is None else _lib
    n = h.mrph_next_best_assignments(_p(e), len(e), n_agents, n_tasks, max_solutions, _p(costs), _p(sol))
    if n < 0:
        raise RuntimeError("mrph_next_best_assignments failed")
    return costs[:n], sol[:n, :n_agents]


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def solve_batch(algo, instances, w=1.0, max_hl=0, max_ll=12000, max_seconds=0.0,
                max_task_assignments=10**9, path_cap=None, max_ll_total=0):
    """instances: objects with dimx, dimy, obstacles [n,2], starts [n,2] and
    goals [n,2] (cbs/ecbs) or potential_goals (list of [k,2]) for cbs_ta.
    Returns a list of dicts (status, cost, makespan, lower_bound, hl_expanded,
    ll_expanded, n_task_assignments, runtime, paths[(x, y, g)...])."""
    instances = list(instances)
    n = len(instances)
    dims = _i32([[i.dimx, i.dimy] for i in instances]).reshape(-1, 2)
    ooff = np.zeros(n + 1, np.int32)
    aoff = np.zeros(n + 1, np.int32)
    for k, i in enumerate(instances):
        ooff[k + 1] = ooff[k] + len(i.obstacles)
        aoff[k + 1] = aoff[k] + len(i.starts)
    obst = _i32(np.concatenate([np.asarray(i.obstacles).reshape(-1, 2) for i in instances])
                if n else [])
    starts = _i32(np.concatenate([i.cell(np.asarray(i.starts)) for i in instances]) if n else [])
    goals = pg_off = pg_cell = None
    total = int(aoff[-1])
    if algo in (CBS_TA, ECBS_TA):
        pg_off = np.zeros(total + 1, np.int32)
        cells = []
        a = 0
        for i in instances:
            for pg in i.potential_goals:
                pg = np.asarray(pg).reshape(-1, 2)
                cells += list(pg[:, 0] + i.dimx * pg[:, 1])
                pg_off[a + 1] = len(cells)
                a += 1
        pg_cell = _i32(cells)
    else:
        goals = _i32(np.concatenate([i.cell(np.asarray(i.goals)) for i in instances]) if n else [])
    if path_cap is None:
        path_cap = max(1024, total * 512)
    status = np.zeros(n, np.int32)
    i64 = lambda: np.zeros(n, np.int64)
    cost, mk, lb, hl, ll, nta = i64(), i64(), i64(), i64(), i64(), i64()
    rt = np.zeros(n, np.float64)
    poff = np.zeros(total + 1, np.int32)
    pcell = np.zeros(path_cap, np.int32)
    pg_ = np.zeros(path_cap, np.int32)
    lib().mrph_set_ll_total(C.c_int64(max_ll_total))
    rc = lib().mrph_solve_batch(algo, n, _p(dims), _p(ooff), _p(obst), _p(aoff), _p(starts),
                                _p(goals), _p(pg_off), _p(pg_cell), C.c_float(w),
                                C.c_int64(max_hl), C.c_int32(max_ll), C.c_double(max_seconds),
                                C.c_int64(max_task_assignments), _p(status), _p(cost), _p(mk),
                                _p(lb), _p(hl), _p(ll), _p(nta), _p(rt), _p(poff), _p(pcell),
                                _p(pg_), C.c_int64(path_cap))
    if rc != 0:
        raise _capi.MrpError(rc, lib().mrph_last_error().decode() or "path buffer too small")
    out = []
    for k, inst in enumerate(instances):
        r = {"status": int(status[k]), "cost": int(cost[k]), "makespan": int(mk[k]),
             "lower_bound": int(lb[k]), "hl_expanded": int(hl[k]), "ll_expanded": int(ll[k]),
             "n_task_assignments": int(nta[k]), "runtime": float(rt[k])}
        if r["status"] == SOLVED:
            # one (x, y, g) array per instance, the agents' paths are views into it
            a0, a1 = int(aoff[k]), int(aoff[k + 1])
            lo, hi = int(poff[a0]), int(poff[a1])
            c = pcell[lo:hi]
            xyg = np.stack([c % inst.dimx, c // inst.dimx, pg_[lo:hi]], 1)
            cut = poff[a0:a1 + 1] - lo
            r["paths"] = [xyg[cut[i]:cut[i + 1]] for i in range(a1 - a0)]
        out.append(r)
    return out


def load_instance_cli_parser(path, ta=False, cap=1 << 16):
    """Parses a YAML file with the C++ reader of the command-line binaries."""
    dims = np.zeros(2, np.int32)
    n_obst = C.c_int32(0)
    obst = np.zeros(cap, np.int32)
    starts = np.zeros(cap, np.int32)
    goals = np.zeros(cap, np.int32)
    pg_off = np.zeros(cap, np.int32)
    pg_cell = np.zeros(cap, np.int32)
    n = lib().mrph_load_instance(path.encode(), int(ta), _p(dims), C.byref(n_obst), _p(obst),
                                 _p(starts), _p(goals), _p(pg_off), _p(pg_cell), cap)
    if n < 0:
        raise _capi.MrpError(n, lib().mrph_last_error().decode())
    d = {"dimx": int(dims[0]), "dimy": int(dims[1]),
         "obstacles": obst[:2 * n_obst.value].reshape(-1, 2).copy(), "starts": starts[:n].copy()}
    if ta:
        d["potential_goals"] = [pg_cell[pg_off[a]:pg_off[a + 1]].copy() for a in range(n)]
    else:
        d["goals"] = goals[:n].copy()
    return d
