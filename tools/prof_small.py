"""Kernel-only timing of the single-tile BFS kernel over all benchmark goals."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, ctypes as C
import libmultirobotplanning_b200 as pkg
capi = pkg.capi; capi.init(0)
for tag in ("32x32", "8x8"):
    sset = pkg.instances.load_set(os.path.join(ROOT, "tests/golden/bench_%s.npz" % tag))
    reps = 40 if tag == "32x32" else 20
    maps = [capi.Map(i.dimx, i.dimy, i.obstacles) for i in sset]
    gm = np.concatenate([np.full(i.n_agents, k, np.int32) for k, i in enumerate(sset)])
    gc = np.concatenate([i.cell(i.goals) for i in sset]).astype(np.int32)
    gm, gc = np.tile(gm, reps), np.tile(gc, reps)
    handles = (C.c_void_p * len(maps))(*[m.handle for m in maps])
    lib = capi.lib()
    for it in range(3):
        fs = C.c_void_p()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        import time
        t0 = time.perf_counter()
        capi.check(lib.mrp_fieldset_create(handles, len(maps), gm.ctypes.data_as(C.c_void_p),
                                           gc.ctypes.data_as(C.c_void_p), len(gc), C.byref(fs)))
        dt = time.perf_counter() - t0
        lib.mrp_fieldset_destroy(fs)
    cells = len(gc) * sset[0].dimx * sset[0].dimy
    print("%s: %d fields (%d maps x goals x %d reps), fieldset_create wall %.3f ms -> %.3g cells/s (incl. alloc + uploads)"
          % (tag, len(gc), len(maps), reps, dt * 1e3, cells / dt))
