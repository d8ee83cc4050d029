"""ctypes binding of libmrp_b200.so — the C ABI declared in include/mrp_b200.h.

This is the only way Python code (tests, bench.py) reaches the CUDA path; the
binding is a 1:1 mirror of the header.  There is no CPU fallback: if the
shared library is missing or no CUDA device is present the calls raise.
"""
import ctypes as C
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
# MRP_B200_LIB: A/B timing of two builds of the same library (tools/), never a fallback
LIB_PATH = os.environ.get("MRP_B200_LIB") or os.path.join(_PKG, "libmrp_b200.so")
INF = 2147483647


class MrpError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("mrp_b200 error %d: %s" % (code, msg))
        self.code = code


class Conflict(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("time", "agent1", "agent2", "type", "x1", "y1", "x2", "y2")]

    def astuple(self):
        return (self.time, self.agent1, self.agent2, self.type,
                self.x1, self.y1, self.x2, self.y2)


class Job(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("map", "start_cell", "goal_cell", "field", "vc_begin", "vc_end",
                 "ec_begin", "ec_end", "table", "self_idx")]


class PathInfo(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("status", "cost", "fmin", "length", "expanded")]


class LowLevelParams(C.Structure):
    _fields_ = [("variant", C.c_int32), ("w", C.c_float),
                ("max_expanded", C.c_int32), ("path_cap", C.c_int32)]


EXPORTS = [
    "mrp_init", "mrp_shutdown", "mrp_device_count", "mrp_last_error",
    "mrp_device_info", "mrp_map_create", "mrp_map_destroy", "mrp_bfs_fields",
    "mrp_bfs_fields_batch", "mrp_bfs_workspace_bytes", "mrp_bfs_fields_dev",
    "mrp_first_conflict", "mrp_count_conflicts", "mrp_conflicts_batch",
    "mrp_focal_counts", "mrp_conflicts_dev", "mrp_decode_conflict",
    "mrp_lowlevel_batch", "mrp_launch_count", "mrp_fieldset_create",
    "mrp_fieldset_read", "mrp_fieldset_destroy", "mrp_lowlevel_batch_fs",
    "mrp_set_lane", "mrp_max_lanes", "mrp_widen_u16", "mrp_widen_u8", "mrp_bfs_d2h_bytes",
    "mrp_comm_unique_id", "mrp_comm_init_rank", "mrp_comm_info", "mrp_comm_destroy",
    "mrp_bfs_allgather_workspace_bytes", "mrp_bfs_fields_allgather_dev", "mrp_comm_last_gather",
    "mrp_conflicts_sharded_dev", "mrp_bfs_fields_packed",
    "mrp_pathpool_create", "mrp_pathpool_destroy", "mrp_pathpool_reserve", "mrp_pathpool_write",
    "mrp_pathpool_read", "mrp_conflicts_batch_pool", "mrp_lowlevel_batch_pool",
    "mrp_pathpool_reserve_states", "mrp_lowlevel_batch_pool_sliced",
    "mrp_free_cell_index", "mrp_bfs_fields_compact", "mrp_bitmap_row_division",
]
COMM_ID_BYTES = 128

_lib = None


def lib():
    """Loads the shared library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise MrpError(-1, "%s not found: build it with "
                           "`python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)" % LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.mrp_last_error.restype = C.c_char_p
        _lib.mrp_device_info.restype = C.c_char_p
        _lib.mrp_bfs_workspace_bytes.restype = C.c_size_t
        _lib.mrp_bfs_workspace_bytes.argtypes = [C.c_void_p, C.c_int]
        _lib.mrp_launch_count.restype = C.c_longlong
        _lib.mrp_bfs_d2h_bytes.restype = C.c_longlong
        _lib.mrp_widen_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                      C.c_int, C.c_int]
        _lib.mrp_widen_u16.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        _lib.mrp_map_destroy.argtypes = [C.c_void_p]
        _lib.mrp_bfs_fields_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int,
                                            C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.mrp_conflicts_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int,
                                           C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p]
        _lib.mrp_decode_conflict.argtypes = [C.c_ulonglong, C.c_int, C.c_int32,
                                             C.c_int32, C.c_void_p]
        _lib.mrp_bfs_allgather_workspace_bytes.restype = C.c_size_t
        _lib.mrp_bfs_allgather_workspace_bytes.argtypes = [C.c_void_p, C.c_int]
        _lib.mrp_bfs_fields_allgather_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_int,
                                                      C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.mrp_conflicts_sharded_dev.argtypes = _lib.mrp_conflicts_dev.argtypes
        _lib.mrp_comm_unique_id.argtypes = [C.c_void_p]
        _lib.mrp_comm_init_rank.argtypes = [C.c_void_p, C.c_int, C.c_int]
    return _lib


def check(rc):
    if rc < 0:
        raise MrpError(rc, lib().mrp_last_error().decode())
    return rc


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def init(device=-1):
    check(lib().mrp_init(device))


def set_lane(lane):
    """Binds the calling thread to a lane of the library (own streams, staging
    buffers and replan workspace); calls from different lanes overlap."""
    check(lib().mrp_set_lane(int(lane)))


def device_info():
    return lib().mrp_device_info().decode()


def launch_count():
    return int(lib().mrp_launch_count())


# ---- multi-GPU (one process per GPU; NCCL behind the C ABI) ----------------------
def comm_unique_id():
    """128 bytes made on rank 0; hand them to the other ranks by any side channel."""
    buf = C.create_string_buffer(COMM_ID_BYTES)
    check(lib().mrp_comm_unique_id(buf))
    return buf.raw


def comm_init_rank(comm_id, n_ranks, rank):
    assert len(comm_id) == COMM_ID_BYTES
    check(lib().mrp_comm_init_rank(C.create_string_buffer(bytes(comm_id), COMM_ID_BYTES), n_ranks, rank))


def comm_info():
    r, n, v = C.c_int(0), C.c_int(1), C.c_int(0)
    have = lib().mrp_comm_info(C.byref(r), C.byref(n), C.byref(v))
    return {"initialised": bool(have), "rank": r.value, "n_ranks": n.value, "nccl_version": v.value}


def comm_destroy():
    check(lib().mrp_comm_destroy())


def comm_last_gather():
    ms, tb, wb, fmt = C.c_double(0), C.c_longlong(0), C.c_longlong(0), C.c_int(0)
    check(lib().mrp_comm_last_gather(C.byref(ms), C.byref(tb), C.byref(wb), C.byref(fmt)))
    return {"collective_ms": ms.value, "timed_bytes_in": tb.value, "wire_bytes_in": wb.value,
            "bytes_per_cell": fmt.value}


class Map:
    """Device-resident bit-packed map (mrp_map)."""

    def __init__(self, dimx, dimy, obst_xy):
        obst = _i32(obst_xy).reshape(-1, 2)
        h = C.c_void_p()
        check(lib().mrp_map_create(dimx, dimy, _p(obst), len(obst), C.byref(h)))
        self.handle = h
        self.dimx, self.dimy = dimx, dimy

    def close(self):
        if self.handle:
            lib().mrp_map_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def workspace_bytes(self, n_goals):
        return int(lib().mrp_bfs_workspace_bytes(self.handle, n_goals))

    def bfs_fields_dev(self, d_goal_cell_ptr, n_goals, d_out_ptr, d_ws_ptr, stream=0):
        check(lib().mrp_bfs_fields_dev(self.handle, d_goal_cell_ptr, n_goals,
                                       d_out_ptr, d_ws_ptr, stream))

    def allgather_workspace_bytes(self, n_goals):
        return int(lib().mrp_bfs_allgather_workspace_bytes(self.handle, n_goals))

    def bfs_fields_allgather_dev(self, d_goal_cell_ptr, n_goals, d_out_ptr, d_ws_ptr, stream=0):
        """Fields of ALL n_goals goals on every rank: computed by goal slice, gathered as detour
        bytes over NCCL, expanded on the device (mrp_bfs_fields_allgather_dev)."""
        check(lib().mrp_bfs_fields_allgather_dev(self.handle, d_goal_cell_ptr, n_goals,
                                                 d_out_ptr, d_ws_ptr, stream))


def bfs_fields(dimx, dimy, obst_xy, goal_xy, out=None):
    obst = _i32(obst_xy).reshape(-1, 2)
    goals = _i32(goal_xy).reshape(-1, 2)
    if out is None:
        out = np.empty((len(goals), dimy * dimx), np.int32)
    check(lib().mrp_bfs_fields(dimx, dimy, _p(obst), len(obst), _p(goals),
                               len(goals), _p(out)))
    return out


def bfs_fields_packed(dimx, dimy, obst_xy, goal_xy, out=None):
    """One detour byte per cell, (distance - Manhattan) / 2, 255 = INF (mrp_bfs_fields_packed).
    Returns (bytes [n][cells], overflowed [n])."""
    obst = _i32(obst_xy).reshape(-1, 2)
    goals = _i32(goal_xy).reshape(-1, 2)
    if out is None:
        out = np.empty((len(goals), dimy * dimx), np.uint8)
    ovf = np.zeros(len(goals), np.int32)
    check(lib().mrp_bfs_fields_packed(dimx, dimy, _p(obst), len(obst), _p(goals), len(goals),
                                      _p(out), _p(ovf)))
    return out, ovf


def free_cell_index(dimx, dimy, obst_xy):
    """(free_bits uint32 [ceil(cells/32)], prefix int32 [ceil(cells/32)+1], n_free): mrp_free_cell_index
    (host only: works without a device)."""
    obst = _i32(obst_xy).reshape(-1, 2)
    nw = (dimx * dimy + 31) // 32
    bits = np.zeros(nw, np.uint32)
    prefix = np.zeros(nw + 1, np.int32)
    n = lib().mrp_free_cell_index(dimx, dimy, _p(obst), len(obst), _p(bits), _p(prefix))
    check(min(n, 0))
    return bits, prefix, int(n)


def bfs_fields_compact(dimx, dimy, obst_xy, goal_xy, n_free, out=None):
    """Detour bytes of the free cells only (mrp_bfs_fields_compact): (bytes [n][n_free], overflowed [n])."""
    obst = _i32(obst_xy).reshape(-1, 2)
    goals = _i32(goal_xy).reshape(-1, 2)
    if out is None:
        out = np.empty((len(goals), n_free), np.uint8)
    ovf = np.zeros(len(goals), np.int32)
    check(lib().mrp_bfs_fields_compact(dimx, dimy, _p(obst), len(obst), _p(goals), len(goals),
                                       _p(out), _p(ovf)))
    return out, ovf


def unpack_compact(compact, free_bits, dimx, dimy, goal_xy):
    """numpy mirror of mrp_compact_value over whole fields: int32 [n][cells]."""
    cells = dimx * dimy
    free = np.unpackbits(free_bits.view(np.uint8), bitorder="little")[:cells].astype(bool)
    full = np.full((len(compact), cells), 255, np.uint8)
    full[:, free] = compact
    out = unpack_field(full, dimx, dimy, goal_xy)
    goals = _i32(goal_xy).reshape(-1, 2)
    out[np.arange(len(goals)), goals[:, 0] + dimx * goals[:, 1]] = 0  # also for a goal on an obstacle
    return out


def unpack_field(packed, dimx, dimy, goal_xy):
    """numpy mirror of mrp_packed_value over whole fields: int32 [n][cells]."""
    goals = _i32(goal_xy).reshape(-1, 2)
    yy, xx = np.mgrid[0:dimy, 0:dimx]
    out = np.empty((len(goals), dimy * dimx), np.int32)
    for k, (gx, gy) in enumerate(goals):
        m = (np.abs(xx - gx) + np.abs(yy - gy)).reshape(-1)
        b = packed[k].astype(np.int32)
        out[k] = np.where(b == 255, INF, 2 * b + m)
    return out


def widen_u8(src, dimx, dimy, goal_cell, threads=4, out=None):
    """Host half of the one-byte field transfer (needs no device):
    src [n_fields, dimy*dimx] uint8 detours -> int32 distances."""
    src = np.ascontiguousarray(src, np.uint8)
    goal_cell = np.ascontiguousarray(goal_cell, np.int32)
    if out is None:
        out = np.empty(src.shape, np.int32)
    check(lib().mrp_widen_u8(_p(src), _p(out), dimx, dimy, _p(goal_cell), len(goal_cell), threads))
    return out


def bfs_d2h_bytes():
    """Device-to-host bytes of the last bfs_fields call."""
    return int(lib().mrp_bfs_d2h_bytes())


def widen_u16(src, threads=4, out=None):
    """Host half of the packed field transfer (needs no device)."""
    src = np.ascontiguousarray(src, np.uint16)
    if out is None:
        out = np.empty(src.shape, np.int32)
    check(lib().mrp_widen_u16(_p(src), _p(out), src.size, threads))
    return out


def bfs_fields_batch(instances):
    """instances: iterable of objects with dimx, dimy, obstacles, goals.
    Returns a list of [n_goals, cells] int32 arrays (one per instance)."""
    instances = list(instances)
    dims = _i32([[i.dimx, i.dimy] for i in instances]).reshape(-1, 2)
    ooff = np.zeros(len(instances) + 1, np.int32)
    goff = np.zeros(len(instances) + 1, np.int32)
    for k, i in enumerate(instances):
        ooff[k + 1] = ooff[k] + len(i.obstacles)
        goff[k + 1] = goff[k] + len(i.goals)
    obst = _i32(np.concatenate([np.asarray(i.obstacles).reshape(-1, 2)
                                for i in instances]) if instances else [])
    goals = _i32(np.concatenate([np.asarray(i.goals).reshape(-1, 2)
                                 for i in instances]) if instances else [])
    sizes = [len(i.goals) * i.dimx * i.dimy for i in instances]
    out = np.empty(int(sum(sizes)), np.int32)
    check(lib().mrp_bfs_fields_batch(len(instances), _p(dims), _p(ooff), _p(obst),
                                     _p(goff), _p(goals), _p(out)))
    res, off = [], 0
    for i, s in zip(instances, sizes):
        res.append(out[off:off + s].reshape(len(i.goals), i.dimx * i.dimy))
        off += s
    return res


def first_conflict(cell, length, dimx, mode):
    cell, length = _i32(cell), _i32(length)
    N, Tpad = cell.shape
    c = Conflict()
    found = check(lib().mrp_first_conflict(_p(cell), _p(length), N, Tpad, dimx,
                                           mode, C.byref(c)))
    return c.astuple() if found else None


def count_conflicts(cell, length, mode=0):
    cell, length = _i32(cell), _i32(length)
    N, Tpad = cell.shape
    n = C.c_int32(0)
    check(lib().mrp_count_conflicts(_p(cell), _p(length), N, Tpad, mode,
                                    C.byref(n)))
    return n.value


def conflicts_batch(cell, length, dimx, mode):
    cell, length = _i32(cell), _i32(length)
    B, N, Tpad = cell.shape
    found = np.zeros(B, np.int32)
    counts = np.zeros(B, np.int32)
    confl = (Conflict * B)()
    check(lib().mrp_conflicts_batch(_p(cell), _p(length), B, N, Tpad, dimx, mode,
                                    _p(found), confl, _p(counts)))
    return [confl[b].astuple() if found[b] else None for b in range(B)], counts


class PathPool:
    """Device rows of paths (mrp_pathpool_*): the caller owns the row numbers."""

    def __init__(self, row_cap):
        self.handle = C.c_void_p()
        self.row_cap = row_cap
        check(lib().mrp_pathpool_create(row_cap, C.byref(self.handle)))

    def close(self):
        if self.handle:
            lib().mrp_pathpool_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reserve(self, n):
        check(lib().mrp_pathpool_reserve(self.handle, n))

    def write(self, slots, cells, length):
        slots, cells, length = _i32(slots), _i32(cells), _i32(length)
        assert cells.shape == (len(slots), self.row_cap)
        check(lib().mrp_pathpool_write(self.handle, _p(slots), len(slots), _p(cells), _p(length)))

    def read(self, slots):
        slots = _i32(slots)
        cells = np.zeros((len(slots), self.row_cap), np.int32)
        length = np.zeros(len(slots), np.int32)
        check(lib().mrp_pathpool_read(self.handle, _p(slots), len(slots), _p(cells), _p(length)))
        return cells, length

    def conflicts_batch(self, table_slots, Tpad, dimx, mode):
        ts = _i32(table_slots)
        B, N = ts.shape
        found = np.zeros(B, np.int32)
        counts = np.zeros(B, np.int32)
        confl = (Conflict * max(B, 1))()
        check(lib().mrp_conflicts_batch_pool(self.handle, _p(ts), B, N, Tpad, dimx, mode,
                                             _p(found), confl, _p(counts)))
        return [confl[b].astuple() if found[b] else None for b in range(B)], counts


def focal_counts(cell, length, self_idx, cand_t, cand_from, cand_to):
    cell, length = _i32(cell), _i32(length)
    N, Tpad = cell.shape
    ct, cf, cto = _i32(cand_t), _i32(cand_from), _i32(cand_to)
    s = np.zeros(len(ct), np.int32)
    tr = np.zeros(len(ct), np.int32)
    check(lib().mrp_focal_counts(_p(cell), _p(length), N, Tpad, self_idx, _p(ct),
                                 _p(cf), _p(cto), len(ct), _p(s), _p(tr)))
    return s, tr


def lowlevel_batch(maps, fields, jobs, variant=0, w=0.0, max_expanded=4000,
                   path_cap=512, tables=None, table_len=None):
    """jobs: list of dicts with keys map, start, goal, field (-1: Manhattan),
    vc [(t, cell)...], ec [(t, from, to)...], table (-1), self.  Returns a list
    of dicts(status, cost, fmin, expanded, cells, g)."""
    n = len(jobs)
    arr = (Job * max(n, 1))()
    vc, ec = [], []
    for k, j in enumerate(jobs):
        a = arr[k]
        a.map, a.start_cell, a.goal_cell = j.get("map", 0), j["start"], j["goal"]
        a.field = j.get("field", -1)
        a.vc_begin = len(vc)
        vc += [list(x) for x in j.get("vc", [])]
        a.vc_end = len(vc)
        a.ec_begin = len(ec)
        ec += [list(x) for x in j.get("ec", [])]
        a.ec_end = len(ec)
        a.table, a.self_idx = j.get("table", -1), j.get("self", 0)
    vc = _i32(vc).reshape(-1, 2)
    ec = _i32(ec).reshape(-1, 3)
    handles = (C.c_void_p * len(maps))(*[m.handle for m in maps])
    nf = 0
    fptr = None
    if fields is not None:
        fields = _i32(fields)
        nf = fields.shape[0]
        fptr = _p(fields)
    nt = N = Tpad = 0
    tptr = lptr = None
    if tables is not None:
        tables, table_len = _i32(tables), _i32(table_len)
        nt, N, Tpad = tables.shape
        tptr, lptr = _p(tables), _p(table_len)
    params = LowLevelParams(variant, w, max_expanded, path_cap)
    info = (PathInfo * max(n, 1))()
    cells = np.zeros((max(n, 1), path_cap), np.int32)
    g = np.zeros((max(n, 1), path_cap), np.int32)
    check(lib().mrp_lowlevel_batch(handles, len(maps), fptr, nf, _p(vc), len(vc),
                                   _p(ec), len(ec), tptr, lptr, nt, N, Tpad, arr, n,
                                   C.byref(params), info, _p(cells), _p(g)))
    out = []
    for k in range(n):
        i = info[k]
        L = i.length if i.status == 0 else 0
        out.append({"status": i.status, "cost": i.cost, "fmin": i.fmin,
                    "expanded": i.expanded, "cells": cells[k, :L].copy(),
                    "g": g[k, :L].copy()})
    return out
