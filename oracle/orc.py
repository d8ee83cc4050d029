"""ctypes binding of the CPU ORACLE (oracle/libmrp_oracle.so).

TEST INFRASTRUCTURE.  Import only from tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never from the product
package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libmrp_oracle.so")
INF = 2147483647
SOLVED, NO_SOLUTION, CAPPED = 0, 1, 2


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("mrp_oracle.cpp", "mrp_oracle.h")]
    if (not force and os.path.exists(_LIB)
            and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in src)):
        return _LIB
    subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _LIB


class Conflict(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("time", "agent1", "agent2", "type", "x1", "y1", "x2", "y2")]

    def astuple(self):
        return (self.time, self.agent1, self.agent2, self.type,
                self.x1, self.y1, self.x2, self.y2)


class Instance(C.Structure):
    _fields_ = [("dimx", C.c_int32), ("dimy", C.c_int32),
                ("n_obst", C.c_int32), ("obst_xy", C.c_void_p),
                ("n_agents", C.c_int32), ("start_xy", C.c_void_p),
                ("goal_xy", C.c_void_p), ("pg_off", C.c_void_p),
                ("pg_xy", C.c_void_p)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("cost", C.c_int64),
                ("makespan", C.c_int64), ("lower_bound", C.c_int64),
                ("hl_expanded", C.c_int64), ("ll_expanded", C.c_int64),
                ("n_task_assignments", C.c_int64), ("runtime_s", C.c_double)]

    def asdict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Caps(C.Structure):
    _fields_ = [("max_hl_expanded", C.c_int64), ("max_ll_expanded", C.c_int64),
                ("max_seconds", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.orc_assignment.restype = C.c_int64
    return _lib


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def floyd_warshall(dimx, dimy, obst_xy):
    obst = _i32(obst_xy).reshape(-1, 2)
    V = dimx * dimy
    out = np.empty((V, V), np.int32)
    lib().orc_floyd_warshall(dimx, dimy, _p(obst), len(obst), _p(out))
    return out


def bfs_fields(dimx, dimy, obst_xy, goal_xy):
    obst = _i32(obst_xy).reshape(-1, 2)
    goals = _i32(goal_xy).reshape(-1, 2)
    out = np.empty((len(goals), dimy * dimx), np.int32)
    lib().orc_bfs_fields(dimx, dimy, _p(obst), len(obst), _p(goals), len(goals),
                         _p(out))
    return out


def first_conflict(cell, length, dimx, mode):
    cell = _i32(cell)
    length = _i32(length)
    N, Tpad = cell.shape
    c = Conflict()
    found = lib().orc_first_conflict(_p(cell), _p(length), N, Tpad, dimx, mode,
                                     C.byref(c))
    return (c.astuple() if found else None)


def count_conflicts(cell, length, mode=0):
    cell = _i32(cell)
    length = _i32(length)
    N, Tpad = cell.shape
    n = C.c_int32(0)
    lib().orc_count_conflicts(_p(cell), _p(length), N, Tpad, mode, C.byref(n))
    return n.value


def focal_counts(cell, length, self_idx, cand_t, cand_from, cand_to):
    cell = _i32(cell)
    length = _i32(length)
    N, Tpad = cell.shape
    ct, cf, cto = _i32(cand_t), _i32(cand_from), _i32(cand_to)
    s = np.zeros(len(ct), np.int32)
    tr = np.zeros(len(ct), np.int32)
    lib().orc_focal_counts(_p(cell), _p(length), N, Tpad, self_idx, _p(ct),
                           _p(cf), _p(cto), len(ct), _p(s), _p(tr))
    return s, tr


def lowlevel(dimx, dimy, obst_xy, variant, start_cell, goal_cell, vc=(), ec=(),
             w=0.0, others=None, others_len=None, self_idx=0,
             max_expanded=2_000_000, path_cap=4096):
    obst = _i32(obst_xy).reshape(-1, 2)
    vc = _i32(vc).reshape(-1, 2)
    ec = _i32(ec).reshape(-1, 3)
    oc = ol = None
    on = ot = 0
    if others is not None:
        oc = _i32(others)
        ol = _i32(others_len)
        on, ot = oc.shape
    cost, fmin, plen = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    exp = C.c_int64(0)
    path = np.zeros((path_cap, 3), np.int32)
    st = lib().orc_lowlevel(dimx, dimy, _p(obst), len(obst), variant,
                            start_cell, goal_cell, _p(vc), len(vc), _p(ec),
                            len(ec), C.c_float(w), _p(oc), _p(ol), on, ot,
                            self_idx, C.c_int64(max_expanded), C.byref(cost),
                            C.byref(fmin), C.byref(exp), _p(path), path_cap,
                            C.byref(plen))
    return {"status": st, "cost": cost.value, "fmin": fmin.value,
            "expanded": exp.value, "path": path[:plen.value].copy()}


def _instance(dimx, dimy, obst_xy, start_xy, goal_xy=None, potential_goals=None):
    keep = []
    obst = _i32(obst_xy).reshape(-1, 2)
    starts = _i32(start_xy).reshape(-1, 2)
    inst = Instance()
    inst.dimx, inst.dimy = dimx, dimy
    inst.n_obst, inst.obst_xy = len(obst), _p(obst)
    inst.n_agents, inst.start_xy = len(starts), _p(starts)
    keep += [obst, starts]
    if goal_xy is not None:
        goals = _i32(goal_xy).reshape(-1, 2)
        inst.goal_xy = _p(goals)
        keep.append(goals)
    if potential_goals is not None:
        off = np.zeros(len(starts) + 1, np.int32)
        flat = []
        for i, g in enumerate(potential_goals):
            off[i + 1] = off[i] + len(g)
            flat += [list(p) for p in g]
        flat = _i32(flat).reshape(-1, 2)
        inst.pg_off, inst.pg_xy = _p(off), _p(flat)
        keep += [off, flat]
    return inst, keep


def _run(fn_name, inst, extra, caps, path_cap):
    n = inst.n_agents
    res = Result()
    off = np.zeros(n + 1, np.int32)
    xyg = np.zeros((path_cap, 3), np.int32)
    c = Caps(*(caps or (0, 0, 0.0)))
    fn = getattr(lib(), fn_name)
    fn(C.byref(inst), *extra, C.byref(c), C.byref(res), _p(off), _p(xyg),
       path_cap)
    out = res.asdict()
    if res.status == SOLVED:
        out["paths"] = [xyg[off[i]:off[i + 1]].copy() for i in range(n)]
    return out


def cbs(dimx, dimy, obst_xy, start_xy, goal_xy, caps=None, path_cap=1 << 16):
    inst, keep = _instance(dimx, dimy, obst_xy, start_xy, goal_xy)
    return _run("orc_cbs", inst, (), caps, path_cap)


def ecbs(dimx, dimy, obst_xy, start_xy, goal_xy, w, caps=None,
         path_cap=1 << 16):
    inst, keep = _instance(dimx, dimy, obst_xy, start_xy, goal_xy)
    return _run("orc_ecbs", inst, (C.c_float(w),), caps, path_cap)


def cbs_ta(dimx, dimy, obst_xy, start_xy, potential_goals,
           max_task_assignments=10**9, caps=None, path_cap=1 << 16):
    inst, keep = _instance(dimx, dimy, obst_xy, start_xy,
                           potential_goals=potential_goals)
    return _run("orc_cbs_ta", inst, (C.c_int64(max_task_assignments),), caps,
                path_cap)


def ecbs_ta(dimx, dimy, obst_xy, start_xy, potential_goals, w, max_task_assignments=10**9,
            caps=None, path_cap=1 << 16):
    inst, keep = _instance(dimx, dimy, obst_xy, start_xy, potential_goals=potential_goals)
    return _run("orc_ecbs_ta", inst, (C.c_float(w), C.c_int64(max_task_assignments)), caps,
                path_cap)


def assignment(edges, n_agents, n_tasks):
    e = np.ascontiguousarray(edges, dtype=np.int64).reshape(-1, 3)
    sol = np.full(max(n_agents, 1), -1, np.int32)
    cost = lib().orc_assignment(_p(e), len(e), n_agents, n_tasks, _p(sol))
    return cost, sol[:n_agents]


def next_best_assignments(edges, n_agents, n_tasks, max_solutions=1000):
    e = np.ascontiguousarray(edges, dtype=np.int64).reshape(-1, 3)
    costs = np.zeros(max_solutions, np.int64)
    sol = np.full((max_solutions, max(n_agents, 1)), -1, np.int32)
    n = lib().orc_next_best_assignments(_p(e), len(e), n_agents, n_tasks,
                                        max_solutions, _p(costs), _p(sol))
    return costs[:n], sol[:n, :n_agents]
