"""Runs the batched host drivers (libmrp_host.so: host/hl_search.hpp) against the HOST EMULATION
of the C ABI (tests/emu/mrp_emu.cpp).  Started as a subprocess by tests/test_host_driver_emu.py:
the emulation is loaded FIRST and globally under the soname libmrp_b200.so, so the driver's
imports bind to it; the process never loads the CUDA library.  TEST INFRASTRUCTURE.

stdin: JSON {"emu": path, "runs": [{"algo", "set", "names" | "ta": [...], "w", "max_hl", "max_ll",
"env": {...}}]}; stdout: JSON list of per-run result lists (paths as [[x, y, g], ...])."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


def main():
    spec = json.load(sys.stdin)
    emu = ctypes.CDLL(spec["emu"], mode=ctypes.RTLD_GLOBAL)
    from libmultirobotplanning_b200 import instances as I
    from libmultirobotplanning_b200 import solver
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
    # its NEEDED libmrp_b200.so is the emulation ("host": another build of the driver, for A/B runs)
    solver._lib = ctypes.CDLL(spec.get("host", solver.LIB_PATH))
    solver._lib.mrph_last_error.restype = ctypes.c_char_p
    assert b"host emulation" in ctypes.cast(emu.mrp_device_info, ctypes.CFUNCTYPE(ctypes.c_char_p))()
    sets = {}

    def load(name):
        if name not in sets:
            sets[name] = {i.name: i for i in I.load_set(os.path.join(ROOT, "tests", "golden", name + ".npz"))}
        return sets[name]

    out = []
    for run in spec["runs"]:
        for k, v in run.get("env", {}).items():
            os.environ[k] = v
        if "ta" in run:
            import make_ref_golden_ta as T
            want = set(run["ta"])
            insts = [j[2] for j in T.selection(list(load("bench_8x8").values()), list(load("bench_32x32").values()))
                     if j[1] == "cbs_ta" and j[0] in want]
            assert len(insts) == len(want)
        elif "fixtures" in run:
            import numpy as np
            fx = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_fixtures.json")))
            insts = []
            for name in run["fixtures"]:
                d = fx[name]
                insts.append(I.Instance(name, d["dimx"], d["dimy"], np.array(d["obstacles"], np.int32).reshape(-1, 2),
                                        np.array(d["starts"], np.int32).reshape(-1, 2),
                                        np.array(d["goals"], np.int32).reshape(-1, 2) if "goals" in d else None,
                                        [np.array(p, np.int32).reshape(-1, 2) for p in d["potentialGoals"]]
                                        if "potentialGoals" in d else None))
        else:
            by = load(run["set"])
            insts = [by[n] for n in run["names"]]
        before = (ctypes.c_long * 9)()
        emu.mrp_emu_counters(before)
        res = solver.solve_batch(run["algo"], insts, w=run.get("w", 1.0), max_hl=run.get("max_hl", 0),
                                 max_ll=run.get("max_ll", 12000), max_ll_total=run.get("max_ll_total", 0),
                                 max_seconds=run.get("max_seconds", 0.0))
        after = (ctypes.c_long * 9)()
        emu.mrp_emu_counters(after)
        for r in res:
            if "paths" in r:
                r["paths"] = [p.tolist() for p in r["paths"]]
        cnt = [a - b for a, b in zip(after, before)]
        cnt[8] &= (1 << 64) - 1  # a sum of hashes modulo 2^64
        out.append({"results": res, "counters": cnt})
        for k in run.get("env", {}):
            os.environ.pop(k, None)
    json.dump(out, sys.stdout)


if __name__ == "__main__":
    main()
