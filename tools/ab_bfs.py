"""A/B timing of builds of libmrp_b200.so on the C5 distance-field step.
usage: python tools/ab_bfs.py spec ...   spec = lib.so|default[@ENV=VAL,...] (each runs in its own process)
Prints ms per step and a checksum of all fields (must agree between builds)."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import libmultirobotplanning_b200 as pkg
    G = int(os.environ.get("AB_GOALS", "1184"))
    capi = pkg.capi
    capi.init(0)
    inst = pkg.instances.synthetic_c5(n_agents=max(G, 64))
    cells = 1024 * 1024
    gc = (inst.goals[:G, 0] + 1024 * inst.goals[:G, 1]).astype(np.int32)
    mp = capi.Map(1024, 1024, inst.obstacles)
    d_goals = torch.from_numpy(gc).cuda()
    d_out = torch.empty((G, cells), dtype=torch.int32, device="cuda")
    ws = torch.empty(max(mp.workspace_bytes(G), 256), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream()
    ts = []
    for r in range(5):
        d_out.fill_(-7)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        mp.bfs_fields_dev(d_goals.data_ptr(), G, d_out.data_ptr(), ws.data_ptr(), s.cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    w = torch.arange(1, cells + 1, device="cuda", dtype=torch.int64) % 1000003
    chk = int(((d_out.to(torch.int64) % 65521) * w).sum().item())
    print("%-40s G=%d  ms: %s  min %.3f  (%.1f us/goal/SM)  chk %d" % (
        os.environ.get("AB_TAG", "default"), G,
        " ".join("%.2f" % t for t in ts), min(ts), min(ts) * 1e3 * 148 / G, chk), flush=True)
    sys.exit(0)

for spec in sys.argv[1:]:
    # spec = lib[@ENV=VAL[,ENV=VAL...]]; lib "default" = the in-tree build
    lib, _, envs = spec.partition("@")
    env = dict(os.environ)
    env["AB_TAG"] = spec
    for kv in filter(None, envs.split(",")):
        k, _, v = kv.partition("=")
        env[k] = v
    if lib != "default":
        env["MRP_B200_LIB"] = os.path.abspath(lib)
    subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=env, check=False)
