"""The product's multi-GPU entry points (NCCL behind the C ABI, csrc/multi.cu)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_comm_symbols_and_unique_id_without_a_gpu():
    """-m 'not gpu': the library exports the multi-GPU calls, binds NCCL at run time and can make a
    communicator id on a box without a device (no compute)."""
    from libmultirobotplanning_b200 import capi
    lib = capi.lib()
    for s in ("mrp_comm_unique_id", "mrp_comm_init_rank", "mrp_comm_info", "mrp_comm_destroy",
              "mrp_bfs_allgather_workspace_bytes", "mrp_bfs_fields_allgather_dev",
              "mrp_comm_last_gather", "mrp_conflicts_sharded_dev"):
        assert hasattr(lib, s), s
    info = capi.comm_info()
    assert info == {"initialised": False, "rank": 0, "n_ranks": 1, "nccl_version": info["nccl_version"]}
    if info["nccl_version"]:
        a, b = capi.comm_unique_id(), capi.comm_unique_id()
        assert len(a) == capi.COMM_ID_BYTES and a != b


@pytest.mark.gpu
def test_one_rank_is_the_plain_call(capi):
    """Without a communicator the gathered call is mrp_bfs_fields_dev and the sharded conflict sweep is
    the all-pairs sweep: tools/multi_gpu_check.py with one rank."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "multi_gpu_check.py"), "--goals", "40"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "rank 0 OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_two_ranks_gather_and_sharded_conflicts(capi):
    """Two GPUs: fields by goal slice + packed all-gather + device expansion, the int32 fall-back on a
    maze, conflicts by agent-pair block + all-reduce: equal to the single-GPU results on both ranks."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    # (one retry on another port: a rendezvous right after another two-rank job on the same box
    # has been seen to fail once)
    for port in ("29517", "29519"):
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                            "--master-addr", "127.0.0.1", "--master-port", port,
                            os.path.join(ROOT, "tools", "multi_gpu_check.py"), "--goals", "300"],
                           capture_output=True, text=True, timeout=900)
        if r.returncode == 0:
            break
    assert r.returncode == 0 and "rank 0 OK" in r.stdout and "rank 1 OK" in r.stdout, \
        r.stdout[-3000:] + r.stderr[-3000:]
