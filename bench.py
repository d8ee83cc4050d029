#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Headline metric (BASELINE.json): BFS heuristic cells/s on config C5 — the
synthetic 1024x1024 grid with 20 % obstacles and 4096 goals;
one "step" = one pass of the distance-field kernel over all goals of the rank
(4096 fields = 17.2 GB of int32 per GPU).  The JSON line also carries the
conflict pair-steps/s of the same config (`also`), the HBM roofline of the
dominant kernel, the CPU baseline (oracle port on the host cores) and the
end-to-end number through the host-pointer C-ABI call.

Multi-GPU: one process per GPU (torchrun), STRONG scaling on the named C5 size:
the 4096 goals are sharded by goal over the ranks and every field is resident on
every GPU at the end of the step (mrp_bfs_fields_allgather_dev: kernels by goal
slice, ncclAllGather of detour bytes in chunks overlapped with the next chunk's
kernel, expansion to int32 on the device — the collective is INSIDE
ms_per_step).  `also.weak_scaling_no_collective` is the variant with 4096 goals
per GPU and no exchange; `also.conflict_pair_blocks_*` the conflict sweep by
agent-pair block + all-reduce; the ECBS batch is sharded by instance.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DIM = 1024
GOALS_PER_GPU = 4096
E2E_GOALS = 1024


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                mx = float(f[1])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(f[0]))
                    for n, v in zip(names, f[3:7]):
                        if v.lower().startswith("active"):
                            reasons.add(n)
            except ValueError:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa(index, local_rank, ranks_on_node):
    """Binds this process (and the page-locked buffers it allocates from now on) to the host cores
    next to its GPU: /sys/bus/pci/devices/<gpu>/local_cpulist, divided among the ranks that share
    that list.  Eight ranks that leave this to chance put most of their staging memory on one
    socket.  Returns a description for the JSON line."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        with open("/sys/bus/pci/devices/%s/local_cpulist" % bdf) as f:
            txt = f.read().strip()
        cpus = []
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus += list(range(int(a), int(b or a) + 1))
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return "no local cpulist"
        # ranks whose GPUs share this list take consecutive slices of it
        sharers = max(1, min(ranks_on_node, int(round(ranks_on_node * len(allowed) / max(1, os.cpu_count())))))
        k = local_rank % sharers
        per = max(1, len(allowed) // sharers)
        mine = allowed[k * per:(k + 1) * per] or allowed
        os.sched_setaffinity(0, mine)
        return "gpu %s: cores %d-%d of %s" % (bdf, mine[0], mine[-1], txt)
    except Exception as e:  # not fatal: the run is just not bound
        return "not bound (%s)" % e


def c5_instance(n_ranks):
    from libmultirobotplanning_b200 import instances
    return instances.synthetic_c5(dim=DIM, n_agents=GOALS_PER_GPU * n_ranks)


def cpu_bfs_rate(inst, goals, threads):
    """Oracle BFS on the host cores (the checker, here as the CPU baseline)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import orc
    orc.build()
    orc.bfs_fields(inst.dimx, inst.dimy, inst.obstacles, goals[:1])  # warm
    chunks = [goals[i::threads] for i in range(threads) if len(goals[i::threads])]
    t0 = time.perf_counter()
    if threads == 1:
        for i in range(0, len(goals), 32):  # 32 fields (134 MB at C5 size) at a time
            orc.bfs_fields(inst.dimx, inst.dimy, inst.obstacles, goals[i:i + 32])
    else:
        with ThreadPoolExecutor(threads) as ex:  # ctypes releases the GIL
            list(ex.map(lambda g: orc.bfs_fields(inst.dimx, inst.dimy, inst.obstacles, g),
                        chunks))
    dt = time.perf_counter() - t0
    return len(goals) * inst.dimx * inst.dimy / dt, dt


def run_reference(args):
    """--impl reference: the reference's CPU path for this metric.  The
    reference itself cannot be built here (Boost / yaml-cpp absent) and its own
    algorithm (Floyd–Warshall over all V = 2^20 cells) is infeasible at this
    size, so the timed code is the oracle's per-goal queue BFS (identical
    output), on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    inst = c5_instance(1)
    per_step = max(cores * 2, 8)
    goals = inst.goals[:per_step]
    for _ in range(min(args.warmup, 1)):
        cpu_bfs_rate(inst, goals[:cores], cores)
    rates, times = [], []
    for _ in range(args.steps):
        r, dt = cpu_bfs_rate(inst, goals, cores)
        rates.append(r)
        times.append(dt)
    value = per_step * DIM * DIM * len(times) / sum(times)
    line = {
        "impl": "reference", "metric": "BFS heuristic cells/s", "value": value,
        "unit": "cells/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": {"workload": "C5 synthetic 1024x1024, 20% obstacles, distance fields by goal",
                   "dim": DIM, "goals_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": "cells/s", "cores": cores, "kind": "port",
                         "sample": "%d goals per step, oracle queue BFS, one thread per core"
                                   % per_step},
        "e2e": {"value": value, "unit": "cells/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def descend_paths(torch, fields, inst, starts_cell, n_agents, max_t):
    """Steepest descent on each agent's goal field with neighbour preference
    Left, Right, Up, Down: the path table of the conflict
    sweep.  fields: device int32 [G][cells], agent i follows field i."""
    dev = fields.device
    cur = torch.as_tensor(starts_cell[:n_agents], device=dev, dtype=torch.int64)
    idx = torch.arange(n_agents, device=dev)
    dist0 = fields[idx, cur].to(torch.int64)
    T = int(dist0.max().item()) + 1
    T = min(T, max_t)
    table = torch.empty((n_agents, T), dtype=torch.int32, device=dev)
    length = (dist0 + 1).clamp(max=T).to(torch.int32)
    x = cur % DIM
    y = cur // DIM
    big = torch.iinfo(torch.int32).max
    for t in range(T):
        table[:, t] = cur.to(torch.int32)
        d = fields[idx, cur]
        nxt = cur.clone()
        done = d == 0
        for dx, dy in ((-1, 0), (1, 0), (0, 1), (0, -1))[::-1]:
            nx, ny = x + dx, y + dy
            ok = (nx >= 0) & (nx < DIM) & (ny >= 0) & (ny < DIM)
            nc = (nx.clamp(0, DIM - 1) + DIM * ny.clamp(0, DIM - 1))
            nd = torch.where(ok, fields[idx, nc], torch.full_like(d, big))
            take = (nd == d - 1) & ~done
            nxt = torch.where(take, nc, nxt)
        cur = nxt
        x = cur % DIM
        y = cur // DIM
    return table.contiguous(), length.contiguous()


ECBS_BATCH = 1000  # instances of one call
# low-level expansions an instance may use in total before it is given up as capped: the solved
# instances of the batch need up to 2.4*10^5; without it one instance in a few thousand runs to the
# high-level cap with 6*10^7 expansions (27 s) and every rank waits for it
LL_TOTAL = 500000


def c3_shard(pkg, s32, rank):
    """The ECBS batch of one rank (config C3, instances sharded across GPUs,
    ECBS_BATCH per GPU).  Rank 0: the 100-agent benchmark files as they are and
    the other files scaled to 100 agents by instances.synthetic_c3.  Rank r > 0:
    the same maps with other agents: every file cut to at most 90 of its agents
    and refilled to 100 with the stateless RNG offset by the rank."""
    I = pkg.instances
    if rank == 0:
        insts = [i for i in s32 if i.n_agents == 100]
        n_files = len(insts)
        insts += [I.synthetic_c3(b, k, 100)
                  for k, b in enumerate(i for i in s32 if i.n_agents != 100)][:ECBS_BATCH - n_files]
        return insts, n_files
    insts = []
    for k, b in enumerate(s32[:ECBS_BATCH]):
        m = min(b.n_agents, 90)
        cut = I.Instance(b.name, b.dimx, b.dimy, b.obstacles, b.starts[:m], b.goals[:m])
        insts.append(I.synthetic_c3(cut, k + 1000 * rank, 100))
    return insts, 0


def search_metrics(pkg, rank=0, world=1, dist=None, dev=None):
    """Instance throughput of the batched searches next to the oracle on one
    host core (same caps).  ECBS (config C3): 1000 instances per GPU on the
    32x32_obst204 maps at 100 agents, w = 1.3, one batch per rank (paths in a
    device pool, replans in resumable slices, instances advance independently),
    no collective on the data path (the ranks only add up their counts); the
    host driver of a rank gets cores / world threads.  CBS: the full 8x8 set
    (config C2) under an expansion cap, rank 0."""
    import torch
    out = {}
    g = os.path.join(ROOT, "tests", "golden")
    s32 = pkg.instances.load_set(os.path.join(g, "bench_32x32.npz"))
    insts, n_files = c3_shard(pkg, s32, rank)
    cap_hl = 2000
    pkg.solver.solve_batch(pkg.solver.ECBS, insts[:2], w=1.3, max_hl=50)  # warm
    if world > 1:
        dist.barrier()
    # two timed runs of the same batch (identical results); the better one counts: the
    # first pays for the growth of the lanes' replan arenas (GB-sized cudaMalloc + clears)
    runs = []
    for _ in range(2):
        t0 = time.perf_counter()
        res = pkg.solver.solve_batch(pkg.solver.ECBS, insts, w=1.3, max_hl=cap_hl, max_seconds=120,
                                     max_ll_total=LL_TOTAL)
        runs.append(time.perf_counter() - t0)
    dt = min(runs)
    ok = [r for r in res if r["status"] == 0]
    n_ok, n_all, dt_max = len(ok), len(insts), dt
    ratio = max(r["cost"] / r["lower_bound"] for r in ok) if ok else 0.0
    if world > 1:
        t = torch.tensor([n_ok, n_all], device=dev, dtype=torch.float64)
        dist.all_reduce(t)
        m = torch.tensor([dt, ratio], device=dev, dtype=torch.float64)
        dist.all_reduce(m, op=dist.ReduceOp.MAX)
        n_ok, n_all, dt_max, ratio = int(t[0].item()), int(t[1].item()), float(m[0].item()), float(m[1].item())
    if rank != 0:
        return out
    from oracle import orc
    s8 = pkg.instances.load_set(os.path.join(g, "bench_8x8.npz"))
    out["ecbs_w1.3_instances_per_s"] = n_ok / dt_max
    out["ecbs_config"] = "32x32_obst204, 100 agents, %d instances per GPU in one batch per rank (paths in a device pool, replans in " \
                         "slices of 256 expansions, instances advance independently; 4 lanes) " \
                         "(rank 0: %d benchmark files + %d scaled by synthetic_c3; other ranks: the same maps, " \
                         "agents redrawn), caps: %d high-level expansions, 12000 expansions per replan, %d low-level " \
                         "expansions per instance; %d rank(s)" % (
                             len(insts), n_files, len(insts) - n_files, cap_hl, LL_TOTAL, world)
    out["ecbs_solved"] = "%d/%d" % (n_ok, n_all)
    out["ecbs_seconds"] = dt_max
    out["ecbs_seconds_runs_rank0"] = runs
    out["ecbs_max_cost_over_lb"] = ratio if n_ok else None
    # every solved instance of the batch through the solution checker (validate.py: starts, goals,
    # unit moves on free cells, no vertex conflict / edge swap inside the reference's loop bound)
    invalid = [(i.name, pkg.validate.validate_paths(i, r["paths"], 0))
               for i, r in zip(insts, res) if r["status"] == 0]
    invalid = [b for b in invalid if b[1]]
    out["ecbs_solutions_checked"] = len(ok)  # this rank's
    out["ecbs_invalid_solutions"] = len(invalid)
    assert not invalid, "invalid ECBS solutions: %s" % invalid[:3]
    out["ecbs_unsolved"] = [i.name for i, r in zip(insts, res) if r["status"] != 0][:40]
    # what the unmodified reference binary does on the instances this path gives up on (its search
    # has no caps: 10 s of wall clock per instance, ten times what the whole batch takes here)
    out["ecbs_unsolved_by_reference_binary"] = reference_on_unsolved(
        "ecbs", [i for i, r in zip(insts, res) if r["status"] != 0][:32], ("-w", "1.3"), 10.0)
    n_cpu = 6
    t0 = time.perf_counter()
    cres = [orc.ecbs(i.dimx, i.dimy, i.obstacles, i.starts, i.goals, 1.3, (cap_hl, 0, 30.0))
            for i in insts[:n_cpu]]
    dt = time.perf_counter() - t0
    out["ecbs_cpu_instances_per_s_1core"] = sum(r["status"] == 0 for r in cres) / dt
    out["ecbs_cpu_sample"] = "%d instances, oracle port" % n_cpu
    ref = reference_binary_rate("ecbs", insts[:n_cpu], ("-w", "1.3"))
    if ref:
        out["ecbs_reference_binary_instances_per_s_1core"] = ref
        out["ecbs_reference_binary"] = ("unmodified example/ecbs.cpp + stand-in Boost/yaml-cpp "
                                        "headers (oracle/_ref), statistics.runtime convention")
    try:  # the same binary on every host core at once (a side measurement: never fatal)
        n_all = max(8, min(32, 2 * (os.cpu_count() or 4)))
        allc = reference_binary_rate_all_cores("ecbs", insts[:n_all], ("-w", "1.3"), timeout=20.0)
        if allc and allc[0]:
            out["ecbs_reference_binary_instances_per_s_all_cores"] = allc[0]
            out["ecbs_reference_binary_all_cores_sample"] = "%d instances, one process per core, %d cores" % (allc[2], allc[1])
    except Exception as e:  # noqa: BLE001
        out["ecbs_reference_binary_all_cores_error"] = repr(e)[:200]
    out["ecbs_cost_gpu_vs_cpu"] = [[a["cost"], b["cost"]] for a, b in zip(res, cres)]
    if world == 1:
        # throughput against batch size: the batch above plus three more of the same kind (the
        # shards ranks 1..3 would get), 4000 instances in one call; the searches of a batch are
        # latency-bound (one dependent chain per instance), so instances/s grows with the batch
        big = list(insts)
        for r in (1001, 1002, 1003):
            big += c3_shard(pkg, s32, r)[0]
        bruns = []
        for _ in range(2):
            t0 = time.perf_counter()
            bres = pkg.solver.solve_batch(pkg.solver.ECBS, big, w=1.3, max_hl=cap_hl, max_seconds=120,
                                          max_ll_total=LL_TOTAL)
            bruns.append(time.perf_counter() - t0)
        bok = sum(r["status"] == 0 for r in bres)
        assert [(r["status"], r["cost"]) for r in bres[:len(insts)]] == [(r["status"], r["cost"]) for r in res]
        out["ecbs_batch4000"] = {"instances": len(big), "solved": bok, "seconds_runs": bruns,
                                 "instances_per_s": bok / min(bruns),
                                 "note": "same call, four times the instances: the first 1000 give the same answers"}
    out.update(c1_datapoint(pkg, s32))
    out.update(c3_scaled(pkg, s32))
    # smem-resident maps: all goals of all 1000 32x32 instances / 2000 8x8 instances
    import torch
    for tag, sset in (("32x32", s32), ("8x8", s8)):
        pkg.capi.bfs_fields_batch(sset[:10])  # warm
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fields = pkg.capi.bfs_fields_batch(sset)
        dt = time.perf_counter() - t0
        if tag == "32x32":
            fields32 = fields
        ncell = sum(f.size for f in fields)
        out["bfs_%s_e2e_cells_per_s" % tag] = ncell / dt
        out["bfs_%s_fields" % tag] = int(sum(len(f) for f in fields))
    # the reference's own algorithm for this precompute: Floyd-Warshall over all
    # cells (shortest_path_heuristic.hpp:47-53), oracle port, one 32x32 map
    t0 = time.perf_counter()
    orc.floyd_warshall(32, 32, s32[0].obstacles)
    out["floyd_warshall_32x32_cpu_seconds_per_map"] = time.perf_counter() - t0
    # ... and the reference's own class doing it (oracle/_ref/sph_fields: shortest_path_heuristic.hpp
    # compiled unmodified, Boost.Graph calls resolved by the stand-in headers), process time included
    sph = os.path.join(ROOT, "oracle", "_ref", "sph_fields")
    if os.path.exists(sph):
        try:  # a side measurement: it must never take the bench line down with it
            import tempfile
            i0 = s32[0]
            txt = "%d %d %d %d\n" % (i0.dimx, i0.dimy, len(i0.obstacles), len(i0.goals))
            txt += "".join("%d %d\n" % (x, y) for x, y in i0.obstacles) + "".join("%d %d\n" % (x, y) for x, y in i0.goals)
            with tempfile.TemporaryDirectory() as td:
                t0 = time.perf_counter()
                raw = subprocess.run([sph], input=txt.encode(), stdout=subprocess.PIPE, check=True, cwd=td,
                                     timeout=120).stdout
                out["floyd_warshall_32x32_reference_class_seconds_per_map"] = time.perf_counter() - t0
            ref_f = np.frombuffer(raw, np.int32).reshape(len(i0.goals), -1)
            out["bfs_32x32_fields_equal_reference_class"] = bool(np.array_equal(ref_f, fields32[0]))
        except Exception as e:  # noqa: BLE001
            out["floyd_warshall_32x32_reference_class_error"] = repr(e)[:200]
    # config C4: CBS-TA with every goal of the instance potential for every agent
    # (all-agents x all-goals distance fields, cost matrix from the fields): the
    # 10- and 20-agent files of the 32x32 set, one batch, next to the oracle
    c4 = [i.with_all_goals_potential() for i in s32 if i.n_agents in (10, 20)]
    cap_hl = 1000
    pkg.solver.solve_batch(pkg.solver.CBS_TA, c4[:2], max_hl=20)  # warm
    # (two runs, the better one counts, as for the ECBS batch: the first grows the device arenas of
    # lanes this process has not used yet)
    c4runs = []
    for _ in range(2):
        t0 = time.perf_counter()
        res = pkg.solver.solve_batch(pkg.solver.CBS_TA, c4, max_hl=cap_hl, max_seconds=120)
        c4runs.append(time.perf_counter() - t0)
    dt = min(c4runs)
    ok = [r for r in res if r["status"] == 0]
    out["cbs_ta_c4_instances_per_s"] = len(ok) / dt
    out["cbs_ta_c4_solved"] = "%d/%d" % (len(ok), len(c4))
    out["cbs_ta_c4_seconds"] = dt
    out["cbs_ta_c4_seconds_runs"] = c4runs
    out["cbs_ta_c4_config"] = "32x32_obst204 files with 10 and 20 agents, potentialGoals = all goals of the " \
                              "instance for every agent, cap %d high-level expansions" % cap_hl
    sub = c4[::25]
    t0 = time.perf_counter()
    cres = [orc.cbs_ta(i.dimx, i.dimy, i.obstacles, i.starts, i.potential_goals, caps=(cap_hl, 0, 20.0))
            for i in sub]
    dt = time.perf_counter() - t0
    out["cbs_ta_c4_cpu_instances_per_s_1core"] = sum(r["status"] == 0 for r in cres) / dt
    out["cbs_ta_c4_cost_mismatches_vs_oracle"] = sum(
        1 for a, b in zip(res[::25], cres) if a["status"] == 0 and b["status"] == 0 and a["cost"] != b["cost"])
    try:  # a side comparison: it must never take the bench line down with it
        out.update(c4_reference_binary(pkg, c4, res))
    except Exception as e:  # noqa: BLE001
        out["cbs_ta_c4_reference_binary_error"] = repr(e)[:200]
    cap_hl = 500
    c2runs = []
    for _ in range(2):  # the first run creates the pools and arenas of the 16 lanes small instances use
        t0 = time.perf_counter()
        res = pkg.solver.solve_batch(pkg.solver.CBS, s8, max_hl=cap_hl, max_seconds=120)
        c2runs.append(time.perf_counter() - t0)
    dt = min(c2runs)
    out["cbs_8x8_solved"] = "%d/%d" % (sum(r["status"] == 0 for r in res), len(s8))
    out["cbs_8x8_seconds"] = dt
    out["cbs_8x8_seconds_runs"] = c2runs
    out["cbs_8x8_hl_expansions_per_s"] = sum(r["hl_expanded"] for r in res) / dt
    out["cbs_8x8_ll_expansions_per_s"] = sum(r["ll_expanded"] for r in res) / dt
    sub = s8[::20]
    t0 = time.perf_counter()
    cres = [orc.cbs(i.dimx, i.dimy, i.obstacles, i.starts, i.goals, (cap_hl, 0, 5.0)) for i in sub]
    dt = time.perf_counter() - t0
    out["cbs_8x8_cpu_hl_expansions_per_s_1core"] = sum(r["hl_expanded"] for r in cres) / dt
    mism = [i.name for i, a, b in zip(sub, res[::20], cres)
            if a["status"] == 0 and b["status"] == 0 and a["cost"] != b["cost"]]
    out["cbs_8x8_cost_mismatches_vs_oracle"] = len(mism)
    return out


def c1_datapoint(pkg, s32):
    """Config C1: ECBS w = 1.3 on map_32by32_obst204_agents10_ex1, one instance (a latency datapoint:
    a 10-agent instance cannot fill a GPU), next to the unmodified reference binary (oracle/_ref/ecbs,
    output format of example/ecbs.cpp:584-617) and the oracle port on one host core."""
    import tempfile
    import yaml
    from oracle import orc
    inst = next(i for i in s32 if i.name == "map_32by32_obst204_agents10_ex1")
    pkg.solver.solve_batch(pkg.solver.ECBS, [inst], w=1.3, max_hl=2000)
    t0 = time.perf_counter()
    r = pkg.solver.solve_batch(pkg.solver.ECBS, [inst], w=1.3, max_hl=2000)[0]
    wall = time.perf_counter() - t0
    assert r["status"] == 0 and pkg.validate.validate_paths(inst, r["paths"], 0) is None
    assert 236 <= r["cost"] <= 306  # optimal sum of costs 236, w = 1.3
    d = {"gpu": {"cost": r["cost"], "makespan": r["makespan"], "lower_bound": r["lower_bound"],
                 "runtime_s": r["runtime"], "wall_s": wall, "highLevelExpanded": r["hl_expanded"],
                 "lowLevelExpanded": r["ll_expanded"]}}
    t0 = time.perf_counter()
    o = orc.ecbs(inst.dimx, inst.dimy, inst.obstacles, inst.starts, inst.goals, 1.3, (2000, 0, 30.0))
    d["oracle_port_1core"] = {"cost": o["cost"], "makespan": o["makespan"], "runtime_s": time.perf_counter() - t0,
                              "highLevelExpanded": o["hl_expanded"], "lowLevelExpanded": o["ll_expanded"]}
    exe = os.path.join(ROOT, "oracle", "_ref", "ecbs")
    if os.path.exists(exe):
        with tempfile.TemporaryDirectory() as td:
            inp, outp = os.path.join(td, "i.yaml"), os.path.join(td, "o.yaml")
            pkg.instances.save_yaml(inst, inp)
            subprocess.run([exe, "-i", inp, "-o", outp, "-w", "1.3"], stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, timeout=60, check=True)
            with open(outp) as f:
                st = yaml.safe_load(f)["statistics"]
        d["reference_binary_1core"] = {"cost": st["cost"], "makespan": st["makespan"], "runtime_s": st["runtime"],
                                       "highLevelExpanded": st["highLevelExpanded"],
                                       "lowLevelExpanded": st["lowLevelExpanded"]}
    return {"c1_ecbs_w1.3_agents10_ex1": d}


def c4_reference_binary(pkg, c4, res):
    """Config C4 next to the UNMODIFIED reference cbs_ta (oracle/_ref/cbs_ta, example/cbs_ta.cpp with
    the reference's own headers against stand-in Boost headers): every answer of the batch against the
    committed answers of that binary (tests/golden/ref_binary_golden_ta.json), and the binary itself
    timed on a few instances on one host core (its time is mostly the Floyd-Warshall precompute over
    all 1024 cells, example/shortest_path_heuristic.hpp:47-53, which its own timer leaves out)."""
    import tempfile
    import yaml
    d = {}
    gpath = os.path.join(ROOT, "tests", "golden", "ref_binary_golden_ta.json")
    if os.path.exists(gpath):
        g = json.load(open(gpath))["cbs_ta"]
        pairs = [(g.get("all/" + i.name), r) for i, r in zip(c4, res)]
        pairs = [(a, b) for a, b in pairs if a is not None and a["solved"] and b["status"] == 0]
        d["cbs_ta_c4_compared_with_reference_binary"] = len(pairs)
        d["cbs_ta_c4_cost_mismatches_vs_reference_binary"] = sum(1 for a, b in pairs if a["cost"] != b["cost"])
    exe = os.path.join(ROOT, "oracle", "_ref", "cbs_ta")
    if os.path.exists(exe):
        walls, inner, same = [], [], 0
        with tempfile.TemporaryDirectory() as td:
            inp, outp = os.path.join(td, "i.yaml"), os.path.join(td, "o.yaml")
            for k in range(0, len(c4), max(1, len(c4) // 4)):
                pkg.instances.save_yaml(c4[k], inp)
                t0 = time.perf_counter()
                subprocess.run([exe, "-i", inp, "-o", outp], stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL, timeout=120, check=True, cwd=td)
                walls.append(time.perf_counter() - t0)
                with open(outp) as f:
                    st = yaml.safe_load(f)["statistics"]
                inner.append(st["runtime"])
                same += int(res[k]["status"] == 0 and st["cost"] == res[k]["cost"])
        d["cbs_ta_c4_reference_binary_1core"] = {
            "instances": len(walls), "same_cost_as_this_path": same,
            "seconds_per_instance_process": sum(walls) / len(walls),
            "seconds_per_instance_search_only": sum(inner) / len(inner)}
    return d


def c3_scaled(pkg, s32, sizes=((120, 100), (160, 50), (200, 30)), seconds=30.0):
    """Config C3 at its stated sizes: the obstacle layouts of the 100-agent files scaled to N agents
    (instances.synthetic_c3), ECBS w = 1.3, one batch of `batch` instances per N under a
    wall-clock cap; solved / cost over lower bound / validity per N.  (Full batches of 100 per N,
    90 s cap: profiles/README.md, round 2.)"""
    base = [i for i in s32 if i.n_agents == 100]
    rows = {}
    for n, batch in sizes:
        insts = [pkg.instances.synthetic_c3(b, k, n) for k, b in enumerate(base[:batch])]
        t0 = time.perf_counter()
        res = pkg.solver.solve_batch(pkg.solver.ECBS, insts, w=1.3, max_hl=2000, max_seconds=seconds)
        dt = time.perf_counter() - t0
        ok = [(i, r) for i, r in zip(insts, res) if r["status"] == 0]
        bad = [i.name for i, r in ok if pkg.validate.validate_paths(i, r["paths"], 0)]
        assert not bad, "invalid ECBS solutions at %d agents: %s" % (n, bad[:3])
        rows[str(n)] = {"solved": "%d/%d" % (len(ok), len(insts)), "seconds": dt,
                        "instances_per_s": len(ok) / dt,
                        "max_cost_over_lb": max((r["cost"] / r["lower_bound"] for _, r in ok), default=None),
                        "invalid_solutions": 0}
    return {"ecbs_c3_scaled": rows,
            "ecbs_c3_scaled_config": "32x32_obst204 layouts of the 100-agent files, agents appended by "
                                     "instances.synthetic_c3, w = 1.3, cap 2000 high-level expansions / "
                                     "%.0f s per batch" % seconds}


def reference_binary_rate(tool, insts, extra, timeout=60.0):
    """instances per second of the reference's own binary (oracle/_ref, built
    from the unmodified sources against stand-in third-party headers), by the
    reference's timing convention: statistics.runtime = search() only."""
    import tempfile
    import yaml
    from libmultirobotplanning_b200 import instances as I
    exe = os.path.join(ROOT, "oracle", "_ref", tool)
    if not os.path.exists(exe):
        return None
    total, solved = 0.0, 0
    with tempfile.TemporaryDirectory() as d:
        for inst in insts:
            inp, outp = os.path.join(d, "i.yaml"), os.path.join(d, "o.yaml")
            I.save_yaml(inst, inp)
            if os.path.exists(outp):
                os.remove(outp)
            try:
                subprocess.run([exe, "-i", inp, "-o", outp, *extra], stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL, timeout=timeout, check=True)
            except Exception:
                total += timeout
                continue
            if os.path.exists(outp):
                with open(outp) as f:
                    total += float(yaml.safe_load(f)["statistics"]["runtime"])
                solved += 1
    return solved / total if total > 0 else None


def reference_binary_rate_all_cores(tool, insts, extra, timeout=60.0):
    """Throughput of the reference's own binary with one process per host core (the reference is
    single-threaded: this is all the host can do with it): `insts` are written out first, then run
    concurrently; instances per second = solved / wall time of the whole set, process start and YAML
    parsing included.  Returns (rate, cores, n) or None."""
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    from libmultirobotplanning_b200 import instances as I
    exe = os.path.join(ROOT, "oracle", "_ref", tool)
    if not os.path.exists(exe) or not insts:
        return None
    allowed = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    cores = max(1, allowed)
    with tempfile.TemporaryDirectory() as d:
        files = []
        for k, inst in enumerate(insts):
            inp = os.path.join(d, "i%d.yaml" % k)
            I.save_yaml(inst, inp)
            files.append((inp, os.path.join(d, "o%d.yaml" % k)))

        def one(io):
            try:
                subprocess.run([exe, "-i", io[0], "-o", io[1], *extra], stdout=subprocess.DEVNULL,
                               stderr=subprocess.DEVNULL, timeout=timeout, check=True, cwd=d)
            except Exception:
                return 0
            return int(os.path.exists(io[1]))
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=cores) as ex:
            solved = sum(ex.map(one, files))
        dt = time.perf_counter() - t0
    return (solved / dt if dt > 0 else None), cores, len(files)


def reference_on_unsolved(tool, insts, extra, wall):
    """Runs oracle/_ref/<tool> on every instance of `insts` with `wall` seconds each, in parallel
    on the host cores; returns how many it solves and their costs."""
    import tempfile
    import yaml
    from concurrent.futures import ThreadPoolExecutor
    from libmultirobotplanning_b200 import instances as I
    exe = os.path.join(ROOT, "oracle", "_ref", tool)
    if not os.path.exists(exe) or not insts:
        return None

    def one(args):
        k, inst, d = args
        inp, outp = os.path.join(d, "i%d.yaml" % k), os.path.join(d, "o%d.yaml" % k)
        I.save_yaml(inst, inp)
        try:
            subprocess.run([exe, "-i", inp, "-o", outp, *extra], stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL, timeout=wall, check=True)
        except Exception:
            return None
        if not os.path.exists(outp):
            return None
        with open(outp) as f:
            st = yaml.safe_load(f)["statistics"]
        return [int(st["cost"]), float(st["runtime"])]

    with tempfile.TemporaryDirectory() as d:
        with ThreadPoolExecutor(max_workers=max(1, min(16, (os.cpu_count() or 2) - 1))) as ex:
            got = list(ex.map(one, [(k, inst, d) for k, inst in enumerate(insts)]))
    return {"instances": len(insts), "wall_cap_s_each": wall,
            "solved_by_reference": sum(g is not None for g in got),
            "cost_and_runtime": {i.name: g for i, g in zip(insts, got) if g is not None}}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import libmultirobotplanning_b200 as pkg
    capi = pkg.capi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    dev = torch.device("cuda", local if world > 1 else 0)
    numa = bind_to_gpu_numa(dev.index, local, world) if world > 1 else "single rank: not bound"
    capi.init(dev.index)

    if world > 1:
        # the product's own communicator (NCCL behind the C ABI, csrc/multi.cu); torch.distributed
        # only carries its 128-byte id and the timing reductions of this script
        idt = torch.zeros(capi.COMM_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            idt = torch.frombuffer(bytearray(capi.comm_unique_id()), dtype=torch.uint8).to(dev)
        dist.broadcast(idt, 0)
        capi.comm_init_rank(idt.cpu().numpy().tobytes(), world, rank)

    inst = c5_instance(1)  # the same 4096-agent instance at every N
    cells = DIM * DIM
    G = GOALS_PER_GPU
    # STRONG scaling, the named C5 size: the G = 4096 goals of the configuration in total, sharded by
    # goal over the ranks, every field resident on every GPU at the end of the step (packed NCCL
    # all-gather + device expansion INSIDE the step: mrp_bfs_fields_allgather_dev).  At N = 1 this is
    # the plain distance-field kernel over 4096 goals.
    goals_xy = inst.goals[:G]
    goal_cells = (goals_xy[:, 0] + DIM * goals_xy[:, 1]).astype(np.int32)
    mp = capi.Map(DIM, DIM, inst.obstacles)
    d_goals = torch.from_numpy(goal_cells).to(dev)
    d_out = torch.empty((G, cells), dtype=torch.int32, device=dev)
    d_ws = torch.empty(max(mp.workspace_bytes(G), mp.allgather_workspace_bytes(G), 256),
                       dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    # the goals of this rank in the weak-scaling variant (4096 per GPU, no collective: `also`)
    from libmultirobotplanning_b200.sharding import shard_range
    g0, g1 = shard_range(world * G, rank, world)
    weak_xy = (inst if world == 1 else c5_instance(world)).goals[g0:g1]
    d_goals_weak = torch.from_numpy((weak_xy[:, 0] + DIM * weak_xy[:, 1]).astype(np.int32)).to(dev)

    def step():
        mp.bfs_fields_allgather_dev(d_goals.data_ptr(), G, d_out.data_ptr(), d_ws.data_ptr(),
                                    stream.cuda_stream)

    def weak_step():
        mp.bfs_fields_dev(d_goals_weak.data_ptr(), G, d_out.data_ptr(), d_ws.data_ptr(),
                          stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(dev.index)
    launches0 = capi.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    t0 = time.time()
    ev[0].record(stream)
    for k in range(args.steps):
        step()
        ev[k + 1].record(stream)
    barrier()
    t1 = time.time()
    launches = capi.launch_count() - launches0
    clocks = sampler.stop(t0, t1)
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = G * cells / (ms_per_step * 1e-3)  # whole job: G goals in total at every N
    gather = capi.comm_last_gather() if world > 1 else None

    # ---- weak scaling, no collective (goals AND their consumers sharded the same way): 4096 goals
    # per GPU; also the timing of the distance-field kernel alone for the roofline at N > 1
    weak_ms = None
    if world > 1:
        for _ in range(2):
            weak_step()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            weak_step()
        e1.record(stream)
        barrier()
        tw = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        weak_ms = float(tw.item())
        step()  # d_out holds the fields of goals 0..G-1 again (the conflict table below walks them)
        barrier()

    # ---- roofline of the dominant kernel (bfs_large_kernel) -------------------
    peak, peak_src = measured_peaks()
    alg_bytes = G * cells * 4 + cells // 8       # 4 B per (goal, cell) + bitmap
    # one launch of the distance-field kernel over 4096 goals: the step itself at N = 1, the
    # no-collective step (4096 goals on this GPU) at N > 1
    kern_ms = float(np.mean(step_ms)) if world == 1 else weak_ms
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        # dram bytes per launch measured by ncu at tj["goals"] goals, scaled per goal
        traffic = tj["dram_bytes_per_goal"] * G
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "kernel": "bfs_queue_kernel", "algorithmic_bytes_per_launch": alg_bytes,
                "kernel_ms": kern_ms}

    # ---- also: conflict sweep of C5 (N = 4096 agents on their goal fields) ----
    also = {}
    table = length = None
    if rank == 0 and not args.skip_conflicts:
        starts_cell = (inst.starts[:, 0] + DIM * inst.starts[:, 1]).astype(np.int64)
        N = G
        table, length = descend_paths(torch, d_out, inst, starts_cell, N, 4096)
        Tpad = table.shape[1]
        d_res = torch.zeros(4, dtype=torch.int64, device=dev)
        lib = capi.lib()

        def cstep():
            capi.check(lib.mrp_conflicts_dev(table.data_ptr(), length.data_ptr(), N, Tpad,
                                             0, 1, 1, d_res.data_ptr(), stream.cuda_stream))
        for _ in range(3):
            cstep()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record(stream)
        for _ in range(reps):
            cstep()
        e1.record(stream)
        torch.cuda.synchronize()
        cms = e0.elapsed_time(e1) / reps
        # first conflict only (what a high-level expansion asks for, cbs.hpp:127):
        # timesteps behind the earliest conflict found so far exit at once
        d_res1 = torch.zeros(4, dtype=torch.int64, device=dev)

        def fstep():
            capi.check(lib.mrp_conflicts_dev(table.data_ptr(), length.data_ptr(), N, Tpad,
                                             0, 1, 0, d_res1.data_ptr(), stream.cuda_stream))
        fstep()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fstep()
        e1.record(stream)
        torch.cuda.synchronize()
        fms = e0.elapsed_time(e1) / reps
        assert int(d_res1[0].item()) == int(d_res[0].item()), "first-conflict key differs"
        max_t = int(length.max().item()) - 1
        pair_steps = N * (N - 1) // 2 * max_t
        res = d_res.cpu().numpy()
        # the answer of the CPU oracle for this very table (tests/golden/make_c5_conflict_golden.py)
        gp = os.path.join(ROOT, "tests", "golden", "c5_conflicts.json")
        golden_ok = None
        if os.path.exists(gp) and world == 1:
            import zlib
            with open(gp) as f:
                gold = json.load(f)
            assert zlib.crc32(table.cpu().numpy().tobytes()) == gold["table_crc32"], "C5 path table differs"
            assert int(res[1]) == gold["count"] and int(np.uint64(res[0])) == gold["first_key"], \
                "C5 conflict sweep differs from the oracle: %s vs %s" % (res[:2], gold)
            golden_ok = True
        also.update({
            "conflict_pair_steps_per_s": pair_steps / (cms * 1e-3),
            "conflict_ms": cms, "conflict_first_only_ms": fms,
            "conflict_agents": N, "conflict_max_t": max_t,
            "conflict_count": int(res[1]),
            "conflict_first_key": int(np.uint64(res[0])) if res[0] != -1 else None,
            "conflict_table_gbps": N * Tpad * 4 / (cms * 1e-3) / 1e9,
            "conflict_equals_oracle_golden": golden_ok,
            "conflict_pair_steps_note": "pair-steps/s is an EQUIVALENT-work figure: N(N-1)/2 * max_t pair tests "
                                        "of the reference's loops divided by the time of the O(N*T) hashed sweep",
        })

    # ---- conflict checks by agent-pair block over the ranks (north_star): every rank
    # holds the table (here: builds it from the gathered fields), sweeps every world-th 64x64 block of
    # agent pairs with the all-pairs kernel, keys meet in an NCCL all-reduce MIN, counts in a SUM
    # (mrp_conflicts_sharded_dev).  On this table one GPU's hashed sweep (above) is faster than any
    # sharding; the pair-block path is the formulation for tables beyond 4096 agents.
    if world > 1 and not args.skip_conflicts:
        if rank != 0:
            starts_cell = (inst.starts[:, 0] + DIM * inst.starts[:, 1]).astype(np.int64)
            table, length = descend_paths(torch, d_out, inst, starts_cell, G, 4096)
        N, Tpad = table.shape
        d_res2 = torch.zeros(4, dtype=torch.int64, device=dev)
        lib = capi.lib()

        def sweep():
            capi.check(lib.mrp_conflicts_sharded_dev(table.data_ptr(), length.data_ptr(), N, Tpad,
                                                     0, 1, 1, d_res2.data_ptr(), stream.cuda_stream))
        for _ in range(2):
            sweep()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 3
        e0.record(stream)
        for _ in range(reps):
            sweep()
        e1.record(stream)
        barrier()
        sms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
        dist.all_reduce(sms, op=dist.ReduceOp.MAX)
        r2 = d_res2.cpu().numpy()
        if rank == 0:
            want_key = also["conflict_first_key"] if also["conflict_first_key"] is not None else -1
            assert int(r2[1]) == also["conflict_count"] and int(np.uint64(r2[0])) == (want_key & (2**64 - 1)), \
                "pair-block sharded conflict sweep differs from the single-GPU sweep"
            also.update({
                "conflict_pair_blocks_ms": float(sms.item()),
                "conflict_pair_blocks_pair_steps_per_s":
                    N * (N - 1) // 2 * also["conflict_max_t"] / (float(sms.item()) * 1e-3),
                "conflict_pair_blocks_how": "all-pairs kernel over every %d-th 64x64 block of agent pairs per rank + "
                                            "NCCL all-reduce MIN(key) / SUM(count), device time, max over ranks; equals "
                                            "the single-GPU result (one GPU: all pairs 7.3 ms, hashed sweep "
                                            "conflict_ms)" % world,
            })
    if table is not None:
        del table, length

    # ---- also: the search metrics (m3): ECBS w=1.3 instances/s, CBS over 8x8 ---
    if not args.skip_search:
        also.update(search_metrics(pkg, rank, world, dist, dev))

    # ---- the collective of the step and the weak-scaling variant ---------------------------
    if world > 1 and rank == 0:
        gbps = gather["timed_bytes_in"] / (gather["collective_ms"] * 1e-3) / 1e9 if gather["collective_ms"] else None
        also.update({
            "strong_scaling": "4096 goals in total; step = distance fields by goal slice + ncclAllGather of detour "
                              "bytes (1 B per cell) in chunks of one wave, overlapped with the next chunk's kernel, "
                              "+ expansion to int32 on the device; every field resident on every GPU",
            "limiting_collective": "ncclAllGather (uint8 detour bytes)",
            "allgather_ms_in_step": gather["collective_ms"],
            "allgather_bytes_in_per_gpu": gather["wire_bytes_in"],
            "allgather_gbps_in_per_gpu": gbps,
            "allgather_frac_of_nvlink5_900gbps": gbps / 900.0 if gbps else None,
            "allgather_bytes_per_cell": gather["bytes_per_cell"],
            "weak_scaling_no_collective": {
                "goals_per_gpu": G, "ms_per_step": weak_ms,
                "cells_per_s": world * G * cells / (weak_ms * 1e-3),
                "note": "goals and their consumers sharded the same way: no exchange (SURVEY §8e)"},
        })

    # ---- e2e: host buffers through the C-ABI calls (H2D + kernels + D2H inside the timed region) ----
    # (i) the packed RESULT mode, mrp_bfs_fields_packed: one detour byte per cell straight into the
    #     caller's page-locked array, read with mrp_packed_value (= getValue); all 4096 goals of the rank;
    # (ii) the int32 contract, mrp_bfs_fields: the same fields as `int` like the reference's matrix
    #     (detour bytes over the bus, expanded by host threads), 1024 goals per step.
    del d_out
    torch.cuda.empty_cache()
    obst = np.ascontiguousarray(inst.obstacles, np.int32)

    def timed(fn, steps):
        fn()  # warm (allocates scratch / staging)
        barrier()
        te0 = time.perf_counter()
        for _ in range(steps):
            fn()
        torch.cuda.synchronize()
        te = (time.perf_counter() - te0) / steps
        t = torch.tensor([te], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    e2e_steps = max(1, min(args.steps, 3))
    Gp = G
    gxy_p = np.ascontiguousarray(weak_xy[:Gp], np.int32)
    h_packed = torch.empty((Gp, cells), dtype=torch.uint8).pin_memory()
    hp_np = h_packed.numpy()
    ovf_box = []

    def packed_call():
        _, ovf = capi.bfs_fields_packed(DIM, DIM, obst, gxy_p, out=hp_np)
        ovf_box.append(int(ovf.sum()))
    te_packed = timed(packed_call, e2e_steps)
    packed_d2h = capi.bfs_d2h_bytes()
    assert ovf_box[-1] == 0, "detours beyond a byte on the C5 map"
    if rank == 0:  # spot-check the packed output against the oracle (checker only)
        from oracle import orc
        probe = [0, 1, Gp // 2, Gp - 1]
        want = orc.bfs_fields(DIM, DIM, inst.obstacles, gxy_p[probe])
        got = capi.unpack_field(hp_np[probe], DIM, DIM, gxy_p[probe])
        assert np.array_equal(got, want), "packed e2e output differs from the oracle"
    del h_packed, hp_np
    # (i') the compact result: the same bytes for the free cells only (mrp_bfs_fields_compact,
    #      mrp_compact_value = getValue): an obstacle cell is INF in every field and the caller has the map
    fbits, fprefix, n_free = capi.free_cell_index(DIM, DIM, obst)
    h_compact = torch.empty((Gp, n_free), dtype=torch.uint8).pin_memory()
    hc_np = h_compact.numpy()

    def compact_call():
        _, ovf = capi.bfs_fields_compact(DIM, DIM, obst, gxy_p, n_free, out=hc_np)
        ovf_box.append(int(ovf.sum()))
    te_compact = timed(compact_call, e2e_steps)
    compact_d2h = capi.bfs_d2h_bytes()
    assert ovf_box[-1] == 0
    if rank == 0:
        from oracle import orc
        probe = [0, 1, Gp // 2, Gp - 1]
        want = orc.bfs_fields(DIM, DIM, inst.obstacles, gxy_p[probe])
        got = capi.unpack_compact(hc_np[probe], fbits, DIM, DIM, gxy_p[probe])
        assert np.array_equal(got, want), "compact e2e output differs from the oracle"
    del h_compact, hc_np

    Ge = min(E2E_GOALS, G)
    h_out = torch.empty((Ge, cells), dtype=torch.int32).pin_memory()
    h_np = h_out.numpy()
    gxy = np.ascontiguousarray(weak_xy[:Ge], np.int32)

    def e2e_rate(steps):
        return timed(lambda: capi.bfs_fields(DIM, DIM, obst, gxy, out=h_np), steps)
    # the int32 transfer (4 B per cell over PCIe) and the uint16 one for comparison,
    # then the default: detour bytes over the bus, expanded to the same int32 array
    # by host threads
    os.environ["MRP_BFS_PACK"] = "0"
    te_plain = e2e_rate(1)
    os.environ.pop("MRP_BFS_PACK")
    os.environ["MRP_BFS_FMT"] = "16"
    te_u16 = e2e_rate(1)
    os.environ.pop("MRP_BFS_FMT")
    h_out.zero_()
    te = e2e_rate(e2e_steps)
    e2e = {"value": world * Gp * cells / te_compact, "unit": "cells/s",
           "h2d_bytes_per_step": int(obst.nbytes + gxy_p.nbytes),
           "d2h_bytes_per_step": int(compact_d2h), "goals_per_step": Gp,
           "ms_per_step": te_compact * 1e3,
           "host_result_bytes_per_step": int(Gp * n_free),
           "api": "mrp_bfs_fields_compact (host pointers, page-locked uint8 output: one detour byte, (distance - "
                  "Manhattan) / 2, 255 = INF, per FREE cell of the map in cell order, read with mrp_compact_value "
                  "= getValue through the free mask and its prefix sums from mrp_free_cell_index; obstacles and "
                  "goals go up, the map is built and every field comes down inside the timed region)",
           "packed_all_cells": {
               "value": world * Gp * cells / te_packed, "unit": "cells/s", "ms_per_step": te_packed * 1e3,
               "d2h_bytes_per_step": int(packed_d2h),
               "api": "mrp_bfs_fields_packed: one detour byte for every cell, obstacles included"},
           "int32_contract": {
               "value": world * Ge * cells / te, "unit": "cells/s", "goals_per_step": Ge,
               "ms_per_step": te * 1e3, "d2h_bytes_per_step": capi.bfs_d2h_bytes(),
               "host_result_bytes_per_step": int(Ge * cells * 4),
               "int32_transfer_cells_per_s": world * Ge * cells / te_plain,
               "uint16_transfer_cells_per_s": world * Ge * cells / te_u16,
               "api": "mrp_bfs_fields (host pointers, pinned int32 output like the reference's int matrix; "
                      "fields cross PCIe as detour bytes and are expanded by host threads; MRP_BFS_FMT=16 "
                      "sends uint16, MRP_BFS_PACK=0 int32)"}}
    # spot-check the e2e output against the oracle (checker only, 4 goals)
    if rank == 0:
        from oracle import orc
        probe = [0, 1, Ge // 2, Ge - 1]
        want = orc.bfs_fields(DIM, DIM, inst.obstacles, gxy[probe])
        assert np.array_equal(h_np[probe], want), "e2e output differs from the oracle"

    # ---- CPU baseline: oracle port, one core, bounded sample -------------------
    cpu = None
    if rank == 0 and world == 1:
        sample = 640  # about 10 s of one host core
        r, dt = cpu_bfs_rate(inst, inst.goals[:sample], 1)
        cpu = {"value": r, "unit": "cells/s", "cores": 1, "kind": "port",
               "sample": "%d goals of the same map, oracle queue BFS (the reference's "
                         "Floyd-Warshall is infeasible at V=2^20), %.1f s" % (sample, dt)}

    if rank == 0:
        line = {
            "metric": "BFS heuristic cells/s", "value": value, "unit": "cells/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": {"workload": "C5 synthetic 1024x1024 grid, 20% obstacles, 4096 goals in total: one BFS "
                                   "distance field per goal, sharded by goal over the GPUs, all fields "
                                   "resident on every GPU after the step (NCCL all-gather inside the step)",
                       "dim": DIM, "goals_total": G, "obstacles": int(len(inst.obstacles)),
                       "l2": "no flush: each step writes 17.2 GB >> 126 MB L2; the 128 KB "
                             "map bitmap is cache-resident by design"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks,
            "gpu_launches": int(launches), "host_binding": numa, "also": also,
            "device": capi.device_info(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        capi.comm_destroy()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-conflicts", action="store_true")
    ap.add_argument("--skip-search", action="store_true")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        # the host drivers of the ranks share the box's cores (read by libgomp when it
        # loads; torchrun presets 1 thread per worker, which would serialise them)
        os.environ["OMP_NUM_THREADS"] = os.environ.get(
            "MRP_BENCH_OMP", str(max(1, (os.cpu_count() or 1) // world)))
        os.environ.setdefault("OMP_WAIT_POLICY", "passive")  # idle workers must not spin on shared cores
        # same for the host threads that expand packed distance fields (mrp_bfs_fields)
        os.environ.setdefault("MRP_WIDEN_THREADS",
                              str(max(1, min(16, (os.cpu_count() or 1) // world))))
        # stdout carries exactly one JSON line: whatever NCCL prints (its banner at NCCL_DEBUG >=
        # VERSION, set by the caller if wanted) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
