"""GPU parity: conflict detection / counting / focal counts vs the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _random_table(rng, N, T, n_cells, ragged=True):
    """Random walks on a line of n_cells cells so conflicts are plentiful."""
    ln = rng.integers(1, T + 1, N) if ragged else np.full(N, T)
    ln[rng.integers(0, N)] = T
    cell = np.full((N, T), -12345, np.int32)
    for i in range(N):
        p = rng.integers(0, n_cells)
        for t in range(ln[i]):
            cell[i, t] = p
            p = int(np.clip(p + rng.integers(-1, 2), 0, n_cells - 1))
    return cell, ln.astype(np.int32)


@pytest.mark.parametrize("N,T,cells", [(2, 2, 3), (3, 9, 4), (10, 45, 40), (65, 70, 300),
                                        (100, 130, 1024), (130, 33, 90), (200, 257, 1024),
                                        (257, 40, 300), (513, 64, 5000), (700, 35, 400),
                                        (1500, 20, 900)])
@pytest.mark.parametrize("mode", [0, 1])
def test_random_tables(capi, orc, N, T, cells, mode):
    rng = np.random.default_rng(N * 7919 + T * 31 + mode)
    for rep in range(3):
        cell, ln = _random_table(rng, N, T, cells)
        assert capi.first_conflict(cell, ln, 32, mode) == orc.first_conflict(cell, ln, 32, mode)
        assert capi.count_conflicts(cell, ln, mode) == orc.count_conflicts(cell, ln, mode)


def test_semantics_cases(capi, orc):
    cases = [
        [[0, 1, 2], [5, 1, 9], [7, 1, 8]],   # 3-agent pile-up: smallest (i,j)
        [[0, 1], [1, 0]],                    # swap
        [[0, 4], [1, 4]],                    # collision only at the final step
        [[0, 1, 2, 3], [3]],                 # clamp to last state
        [[0, 1, 2, 3, 4], [3]],
        [[2, 2, 2], [2, 2, 2]],              # resting pair: vertex + "edge"
        [[0], [1]],                          # length-1 paths, no conflict
        [[3], [3]],                          # length-1 paths on one cell
        [[0, 1, 5], [2, 1, 0], [1, 0, 1]],   # vertex and edge at the same t
    ]
    for paths in cases:
        T = max(len(p) for p in paths)
        cell = np.full((len(paths), T), -99, np.int32)
        ln = np.array([len(p) for p in paths], np.int32)
        for k, p in enumerate(paths):
            cell[k, :len(p)] = p
        for mode in (0, 1):
            assert capi.first_conflict(cell, ln, 10, mode) == orc.first_conflict(cell, ln, 10, mode), (paths, mode)
            assert capi.count_conflicts(cell, ln, mode) == orc.count_conflicts(cell, ln, mode), (paths, mode)


def test_conflict_free_and_degenerate(capi):
    cell = np.arange(12, dtype=np.int32).reshape(4, 3)
    ln = np.full(4, 3, np.int32)
    assert capi.first_conflict(cell, ln, 4, 0) is None
    assert capi.count_conflicts(cell, ln, 0) == 0
    assert capi.first_conflict(cell[:1], ln[:1], 4, 0) is None   # one agent
    with pytest.raises(capi.MrpError):
        capi.first_conflict(cell, np.array([3, 3, 3, 4]), 4, 0)  # len > Tpad
    with pytest.raises(capi.MrpError):
        capi.first_conflict(cell, ln, 4, 2)                      # bad mode


def test_batch(capi, orc):
    rng = np.random.default_rng(3)
    B, N, T = 37, 20, 40
    cells, lens = [], []
    for b in range(B):
        c, l = _random_table(rng, N, T, 60 if b % 3 else 100000)
        cells.append(c)
        lens.append(l)
    cells, lens = np.stack(cells), np.stack(lens)
    for mode in (0, 1):
        confl, counts = capi.conflicts_batch(cells, lens, 32, mode)
        for b in range(B):
            assert confl[b] == orc.first_conflict(cells[b], lens[b], 32, mode)
            assert counts[b] == orc.count_conflicts(cells[b], lens[b], mode)


def test_focal_counts(capi, orc):
    rng = np.random.default_rng(11)
    N, T = 60, 50
    cell, ln = _random_table(rng, N, T, 30)
    ln[5] = 0  # an agent that has no path yet (ecbs.cpp:287)
    n = 500
    ct = rng.integers(0, T + 5, n)
    cf = rng.integers(0, 30, n)
    cto = np.clip(cf + rng.integers(-1, 2, n), 0, 29)
    for self_idx in (0, 5, 17):
        s, tr = capi.focal_counts(cell, ln, self_idx, ct, cf, cto)
        s2, tr2 = orc.focal_counts(cell, ln, self_idx, ct, cf, cto)
        assert np.array_equal(s, s2) and np.array_equal(tr, tr2)
    assert tr2.sum() > 0 and s2.sum() > 0


def test_large_table_properties(capi, orc):
    """C5-sized table (N=4096 agents): the count must be invariant under a
    permutation of the agents and the first conflict must map through it."""
    rng = np.random.default_rng(2)
    N, T = 4096, 96
    cell, ln = _random_table(rng, N, T, 2_000_000, ragged=False)
    # plant a few conflicts
    cell[100, 50:] = cell[3000, 50:]
    cell[7, 20], cell[7, 21] = 5, 6
    cell[4000, 20], cell[4000, 21] = 6, 5
    c0 = capi.count_conflicts(cell, ln, 0)
    f0 = capi.first_conflict(cell, ln, 1024, 0)
    assert f0 is not None and c0 >= 2
    perm = rng.permutation(N)
    c1 = capi.count_conflicts(cell[perm], ln[perm], 0)
    assert c0 == c1
    sub = np.r_[0:64, 100, 3000, 4000]
    assert capi.count_conflicts(cell[sub], ln[sub], 0) == orc.count_conflicts(cell[sub], ln[sub], 0)
    assert capi.first_conflict(cell[sub], ln[sub], 1024, 0) == orc.first_conflict(cell[sub], ln[sub], 1024, 0)


def test_hashed_path_equals_all_pairs(capi, orc, monkeypatch):
    """N in (256, 4096] runs the per-timestep hashing kernels; they must agree
    with the all-pairs kernel (forced through MRP_CONFLICTS_ALLPAIRS) and, on a
    size the oracle can do, with the oracle — including dense pile-ups, resting
    agents and agents without a path."""
    rng = np.random.default_rng(9)
    for N, T, cells in ((300, 50, 200), (1024, 60, 3000), (4096, 40, 100000)):
        cell, ln = _random_table(rng, N, T, cells)
        ln[3] = 0
        cell[10, :], ln[10] = cell[11, 0], T        # resting on agent 11's start
        cell[11, :], ln[11] = cell[11, 0], T
        cell[12, :], ln[12] = cell[11, 0], T
        for mode in (0, 1):
            f_h = capi.first_conflict(cell, ln, 32, mode)
            c_h = capi.count_conflicts(cell, ln, mode)
            monkeypatch.setenv("MRP_CONFLICTS_ALLPAIRS", "1")
            f_p = capi.first_conflict(cell, ln, 32, mode)
            c_p = capi.count_conflicts(cell, ln, mode)
            monkeypatch.delenv("MRP_CONFLICTS_ALLPAIRS")
            assert (f_h, c_h) == (f_p, c_p), (N, mode)
            # the default is the sieve kernel (+ the single-table kernel for dense
            # timesteps); the single-table kernel on its own
            for var in ("MRP_CONFLICTS_HASH2",):
                monkeypatch.setenv(var, "1")
                assert (capi.first_conflict(cell, ln, 32, mode),
                        capi.count_conflicts(cell, ln, mode)) == (f_p, c_p), (N, mode, var)
                monkeypatch.delenv(var)
            if N <= 1024:
                assert f_h == orc.first_conflict(cell, ln, 32, mode)
                assert c_h == orc.count_conflicts(cell, ln, mode)


def test_sieve_sparse_and_dense_timesteps(capi, orc, monkeypatch):
    """One table with sparse timesteps (a handful of candidates: the sieve
    kernel's own exact table), timesteps where most agents share cells (handed
    over to the single-table kernel) and agents without a path."""
    rng = np.random.default_rng(21)
    N, T, cells = 2048, 12, 1 << 20
    cell = rng.integers(0, cells, (N, T)).astype(np.int32)
    cell[:, 5] = rng.integers(0, 700, N)          # a dense timestep
    cell[:, 6] = cell[:, 5]                       # everybody rests
    cell[:, 7] = rng.integers(0, 700, N)
    cell[100, 2], cell[101, 2] = 77, 77           # a lone vertex conflict
    cell[200, 9], cell[200, 10] = 500000, 500001  # a lone swap
    cell[201, 9], cell[201, 10] = 500001, 500000
    cell[300, 3], cell[300, 4] = cells - 1, cells - 2   # the last cell of the map
    cell[301, 3], cell[301, 4] = cells - 2, cells - 1
    ln = np.full(N, T, np.int32)
    ln[7] = 0
    ln[8] = 3
    for mode in (0, 1):
        want = (orc.first_conflict(cell, ln, 1 << 10, mode), orc.count_conflicts(cell, ln, mode))
        assert (capi.first_conflict(cell, ln, 1 << 10, mode), capi.count_conflicts(cell, ln, mode)) == want
    # without the dense timesteps the first conflict is the lone one
    cell[:, 5:8] = rng.integers(0, cells, (N, 3))
    f = capi.first_conflict(cell, ln, 1 << 10, 0)
    assert f == orc.first_conflict(cell, ln, 1 << 10, 0)
    assert capi.count_conflicts(cell, ln, 0) == orc.count_conflicts(cell, ln, 0)


def test_first_conflict_windows(capi, orc):
    """A first-conflict-only call sweeps the first 128 timesteps, then the rest:
    conflicts on both sides of the window boundary (and of other multiples of
    64), in the last step, none at all, and stale hand-over flags from a dense
    table swept just before."""
    N, T, cells = 600, 1500, 1 << 20
    base = np.zeros((N, T), np.int32)
    tt = np.arange(T)
    for i in range(N):
        base[i] = 1000 + 8 * i + (tt + i) % 2      # everybody shuttles between two private cells
    ln = np.full(N, T, np.int32)
    rng = np.random.default_rng(5)
    dense = rng.integers(0, 300, (N, T)).astype(np.int32)   # every timestep is handed over
    assert capi.first_conflict(base, ln, 1024, 0) is None
    assert capi.count_conflicts(base, ln, 0) == 0
    for k, tc in enumerate([0, 63, 64, 126, 127, 128, 129, 320, 1344, T - 2]):
        cell = base.copy()
        i, j = 17 + k, 400 + 3 * k
        if k % 2 == 0:      # vertex conflict at tc
            cell[j, tc] = cell[i, tc]
        else:               # swap between tc and tc + 1
            cell[j, tc], cell[j, tc + 1] = cell[i, tc + 1], cell[i, tc]
        if k == 4:
            capi.count_conflicts(dense, ln, 0)
            assert capi.first_conflict(dense, ln, 1024, 0) == orc.first_conflict(dense, ln, 1024, 0)
        want = orc.first_conflict(cell, ln, 1024, 0)
        assert want is not None and want[0] == tc
        assert capi.first_conflict(cell, ln, 1024, 0) == want, tc
        # a later conflict does not change the answer
        cell[5, min(tc + 70, T - 1)] = cell[6, min(tc + 70, T - 1)]
        assert capi.first_conflict(cell, ln, 1024, 0) == want, tc
    assert capi.first_conflict(base, ln, 1024, 1) is None


def test_hashed_path_stress_high_load(capi, orc):
    """The hashed tables run at a load factor of ~0.5 when N approaches 4096;
    an early version lost table updates there (fire-and-forget shared-memory
    reductions racing the barrier).  Repeat on fresh random moves."""
    rng = np.random.default_rng(77)
    for trial in range(25):
        N, cells = ((4000, 30000), (4096, 50000), (3900, 8000))[trial % 3]
        a = rng.integers(0, cells, N).astype(np.int32)
        b = np.clip(a + rng.integers(-1, 2, N), 0, cells - 1).astype(np.int32)
        cell = np.stack([a, b], 1).astype(np.int32)
        ln = np.full(N, 2, np.int32)
        want_c = orc.count_conflicts(cell, ln, 0)
        want_f = orc.first_conflict(cell, ln, 1000, 0)
        for _ in range(3):
            assert capi.count_conflicts(cell, ln, 0) == want_c
        assert capi.first_conflict(cell, ln, 1000, 0) == want_f


def test_c5_table_equals_the_oracle_golden(capi):
    """The full config-C5 table (4096 agents walking down their goal fields on the synthetic
    1024x1024 map, max_t = 1908): the table built on the GPU is byte-identical to the one the CPU
    oracle built (CRC), and count / first-conflict key of every GPU sweep equal the oracle's
    all-pairs answer (tests/golden/c5_conflicts.json, made by tests/golden/make_c5_conflict_golden.py;
    reference loops: example/cbs.cpp:335-386, example/ecbs.cpp:315-350)."""
    import json
    import os
    import sys
    import zlib
    import torch
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with open(os.path.join(root, "tests", "golden", "c5_conflicts.json")) as f:
        gold = json.load(f)
    sys.path.insert(0, root)
    import bench
    import libmultirobotplanning_b200 as pkg
    N, DIM = gold["N"], 1024
    inst = pkg.instances.synthetic_c5(dim=DIM, n_agents=N)
    gc = (inst.goals[:N, 0] + DIM * inst.goals[:N, 1]).astype(np.int32)
    mp = capi.Map(DIM, DIM, inst.obstacles)
    dev = torch.device("cuda", 0)
    d_goals = torch.from_numpy(gc).to(dev)
    d_out = torch.empty((N, DIM * DIM), dtype=torch.int32, device=dev)
    ws = torch.empty(max(mp.workspace_bytes(N), 256), dtype=torch.uint8, device=dev)
    mp.bfs_fields_dev(d_goals.data_ptr(), N, d_out.data_ptr(), ws.data_ptr(), 0)
    torch.cuda.synchronize()
    starts = (inst.starts[:, 0] + DIM * inst.starts[:, 1]).astype(np.int64)
    table, length = bench.descend_paths(torch, d_out, inst, starts, N, 4096)
    del d_out
    tab = table.cpu().numpy()
    ln = length.cpu().numpy()
    assert tab.shape == (N, gold["T"])
    assert zlib.crc32(tab.tobytes()) == gold["table_crc32"]
    assert zlib.crc32(ln.tobytes()) == gold["length_crc32"]
    lib = capi.lib()
    for first, count in ((1, 1), (1, 0), (0, 1)):
        res = torch.zeros(4, dtype=torch.int64, device=dev)
        capi.check(lib.mrp_conflicts_dev(table.data_ptr(), length.data_ptr(), N, tab.shape[1], 0,
                                         first, count, res.data_ptr(), 0))
        torch.cuda.synchronize()
        r = res.cpu().numpy()
        if first:
            assert int(np.uint64(r[0])) == gold["first_key"]
        if count:
            assert int(r[1]) == gold["count"]
    # and through the host-pointer entry points
    assert capi.count_conflicts(tab, ln, 0) == gold["count"]
    assert list(capi.first_conflict(tab, ln, DIM, 0)) == gold["first_conflict"]


def test_path_pool_rows_and_conflicts(capi, orc):
    """mrp_pathpool_write / _read round trip over two chunks of rows, and
    mrp_conflicts_batch_pool (tables gathered on the device from row numbers, -1 = agent
    without a path) against mrp_conflicts_batch on the same tables."""
    rng = np.random.default_rng(77)
    cap, dimx = 48, 16
    pool = capi.PathPool(cap)
    n = 400
    slots = rng.choice(40000, n, replace=False).astype(np.int32)  # beyond the first 32768-row chunk
    pool.reserve(int(slots.max()) + 1)
    length = rng.integers(1, cap + 1, n).astype(np.int32)
    cells = rng.integers(0, dimx * dimx, (n, cap)).astype(np.int32)
    # random walks so that conflicts are real ones
    for i in range(n):
        c = int(cells[i, 0])
        for t in range(1, cap):
            x, y = c % dimx, c // dimx
            dx, dy = [(0, 0), (1, 0), (-1, 0), (0, 1), (0, -1)][int(rng.integers(0, 5))]
            x, y = min(max(x + dx, 0), dimx - 1), min(max(y + dy, 0), dimx - 1)
            c = x + dimx * y
            cells[i, t] = c
    pool.write(slots, cells, length)
    back, blen = pool.read(slots[::-1].copy())
    assert np.array_equal(blen, length[::-1])
    for k in range(n):
        L = length[n - 1 - k]
        assert np.array_equal(back[k, :L], cells[n - 1 - k, :L])
    B, N = 30, 12
    pick = rng.integers(0, n, (B, N))
    pick[3, 5] = -1
    pick[7, 0] = -1
    table_slots = np.where(pick >= 0, slots[np.maximum(pick, 0)], -1).astype(np.int32)
    Tpad = int(length.max())
    dense = np.zeros((B, N, Tpad), np.int32)
    dlen = np.zeros((B, N), np.int32)
    for b in range(B):
        for a in range(N):
            if pick[b, a] >= 0:
                L = length[pick[b, a]]
                dense[b, a, :L] = cells[pick[b, a], :L]
                dlen[b, a] = L
    for mode in (0, 1):
        ref_c, ref_n = capi.conflicts_batch(dense, dlen, dimx, mode)
        got_c, got_n = pool.conflicts_batch(table_slots, Tpad, dimx, mode)
        assert got_c == ref_c
        assert np.array_equal(got_n, ref_n)
        assert sum(c is not None for c in ref_c) >= B // 2
    with pytest.raises(capi.MrpError):
        pool.read(np.array([10 ** 6], np.int32))
    pool.close()


def test_hashed_sweep_with_parked_agents(capi, orc, monkeypatch):
    """Paths of very different lengths on a small map: agents that walk over, wait on and
    park on parked agents, several paths ending on one cell, paths of one state and agents
    without a path, a pile-up that overflows the sieve's exact table.  The hashed sweep
    (256 < N <= 4096) must equal the all-pairs kernel and the oracle, both loop bounds,
    first conflict and counts.  (Written for a sweep over the moving agents only — sorted by
    path length, parked agents in a static table — which passed it but was slower than the
    full sieve on the C5 table, 0.111 against 0.054 ms: profiles/README.md.)"""
    rng = np.random.default_rng(123)
    for N, T, cells, dimx in ((300, 90, 400, 20), (700, 60, 900, 30), (3000, 48, 4096, 64)):
        dimy = cells // dimx
        cell = np.zeros((N, T), np.int32)
        ln = rng.integers(1, T + 1, N).astype(np.int32)
        ln[rng.integers(0, N, N // 10)] = 1      # parked from the start
        ln[rng.integers(0, N, N // 20)] = T      # move until the end
        ln[5] = 0
        ln[6] = 0
        for i in range(N):                       # random walks (with waits) on a small map: real conflicts
            c = int(rng.integers(0, cells))
            for t in range(T):
                cell[i, t] = c
                x, y = c % dimx, c // dimx
                dx, dy = [(0, 0), (0, 0), (1, 0), (-1, 0), (0, 1), (0, -1)][int(rng.integers(0, 6))]
                x, y = min(max(x + dx, 0), dimx - 1), min(max(y + dy, 0), dimy - 1)
                c = x + dimx * y
        # several paths ending on one cell at different times, one of them waiting there
        goal = int(cell[20, ln[20] - 1])
        for k, i in enumerate((21, 22, 23)):
            ln[i] = max(2, min(T, ln[20] + 3 * k - 2))
            cell[i, ln[i] - 1] = goal
            cell[i, ln[i] - 2] = goal if k == 1 else cell[i, ln[i] - 2]
        if N >= 3000:
            cell[1000:2400, 7] = rng.integers(0, 40, 1400)   # a pile-up among moving agents
            ln[1000:2400] = np.maximum(ln[1000:2400], 12)
        for mode in (0, 1):
            got = (capi.first_conflict(cell, ln, dimx, mode), capi.count_conflicts(cell, ln, mode))
            monkeypatch.setenv("MRP_CONFLICTS_HASH2", "1")
            full = (capi.first_conflict(cell, ln, dimx, mode), capi.count_conflicts(cell, ln, mode))
            monkeypatch.delenv("MRP_CONFLICTS_HASH2")
            monkeypatch.setenv("MRP_CONFLICTS_ALLPAIRS", "1")
            pairs = (capi.first_conflict(cell, ln, dimx, mode), capi.count_conflicts(cell, ln, mode))
            monkeypatch.delenv("MRP_CONFLICTS_ALLPAIRS")
            assert full == pairs, (N, mode)
            assert got == pairs, (N, mode, got, pairs)
            if N <= 700:
                assert got == (orc.first_conflict(cell, ln, dimx, mode), orc.count_conflicts(cell, ln, mode))
            assert got[1] > 50


def test_conflicts_equal_reference_environment(capi):
    """The CUDA conflict kernels against the reference's OWN Environment methods (the example
    files included unmodified into oracle/_ref/env_probe_*; answers committed by
    tests/golden/make_env_golden.py): first conflict under both loop bounds
    (example/cbs.cpp:335-386, cbs_ta.cpp:369-420), focalHeuristic's count (ecbs.cpp:315-350) and
    focalState / focalTransition counts (ecbs.cpp:282-312) on 40 seeded tables."""
    import json
    import os
    import sys
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if gdir not in sys.path:
        sys.path.insert(0, gdir)
    import make_env_golden as E
    g = json.load(open(os.path.join(gdir, "env_probe_golden.json")))
    for k, (dimx, cell, ln, qs) in enumerate(E.tables()):
        for mode, key in ((0, "first_mode0"), (1, "first_mode1")):
            got = capi.first_conflict(cell, ln, dimx, mode)
            assert (list(got) if got else None) == g[key][k], (k, mode)
        assert capi.count_conflicts(cell, ln, 0) == g["count"][k], k
        for q, want in zip(qs, g["focal_queries"][k]):
            s, tr = capi.focal_counts(cell, ln, q[0], [q[1]], [q[2]], [q[3]])
            assert [int(s[0]), int(tr[0])] == want, (k, q)
