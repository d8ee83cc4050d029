import os
"""GPU parity: distance fields through the C ABI vs the CPU oracle (bit-exact)."""
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rand_map(rng, dimx, dimy, density):
    blocked = rng.random((dimy, dimx)) < density
    ys, xs = np.nonzero(blocked)
    return np.stack([xs, ys], 1).astype(np.int32)


def test_anchor_field(capi, set32):
    inst = next(i for i in set32 if i.name == "map_32by32_obst204_agents10_ex1")
    f = capi.bfs_fields(32, 32, inst.obstacles, inst.goals)
    assert zlib.crc32(f[0].astype("<i4").tobytes()) == 0x2C73F098


def test_all_benchmark_instances_batch(capi, orc, set8, set32):
    for sset in (set8, set32):
        got = capi.bfs_fields_batch(sset)
        for inst, g in zip(sset, got):
            want = orc.bfs_fields(inst.dimx, inst.dimy, inst.obstacles, inst.goals)
            assert np.array_equal(g, want), inst.name


def test_golden_crcs(capi, set8, set32, oracle_golden):
    by = {i.name: i for i in set8 + set32}
    for name, g in oracle_golden["bfs_fields"].items():
        i = by[name]
        f = capi.bfs_fields(i.dimx, i.dimy, i.obstacles, i.goals)
        assert zlib.crc32(np.ascontiguousarray(f, "<i4").tobytes()) == g["crc32"], name


@pytest.mark.parametrize("dimx,dimy,density", [
    (1, 1, 0.0), (1, 7, 0.0), (9, 1, 0.2), (5, 2, 0.3), (8, 8, 0.2), (31, 17, 0.25),
    (32, 32, 0.2), (33, 32, 0.2), (32, 33, 0.2), (40, 50, 0.3), (64, 64, 0.2),
    (100, 37, 0.35), (129, 65, 0.2), (257, 300, 0.25), (1100, 70, 0.2),
    (70, 1100, 0.2),
    # narrow and more than one tile high (dimx <= 30): used to get one bitmap word per row, for which the
    # queue kernel's multiply-shift division has no constant (found by test_fields_equal_reference_class)
    (17, 45, 0.0), (30, 33, 0.2), (8, 100, 0.3), (3, 64, 0.0), (29, 1000, 0.2), (1, 40, 0.0),
    # the queue kernel's compile-time instances (widths 256, 512, 1024, 2048)
    (256, 90, 0.2), (512, 300, 0.25), (1024, 130, 0.2), (2048, 75, 0.3)])
def test_random_maps(capi, orc, dimx, dimy, density):
    rng = np.random.default_rng(dimx * 1000 + dimy)
    obst = _rand_map(rng, dimx, dimy, density)
    n = min(dimx * dimy, 12)
    cells = rng.choice(dimx * dimy, n, replace=False)
    goals = np.stack([cells % dimx, cells // dimx], 1)  # may sit on obstacles
    got = capi.bfs_fields(dimx, dimy, obst, goals)
    want = orc.bfs_fields(dimx, dimy, obst, goals)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("env", [{"MRP_BFS_QCAP": "32"}, {"MRP_BFS_TILES": "1"},
                                 {"MRP_BFS_QCAP": "64", "MRP_BFS_THREADS": "64"}])
def test_queue_overflow_and_tiled_fallback(capi, orc, env):
    """The queue kernel hands goals whose wavefront outgrows its shared-memory
    queues to the tiled kernel; MRP_BFS_TILES forces the tiled kernel."""
    rng = np.random.default_rng(77)
    try:
        os.environ.update(env)
        for dimx, dimy, density in [(257, 300, 0.25), (100, 37, 0.1), (40, 50, 0.3)]:
            obst = _rand_map(rng, dimx, dimy, density)
            cells = rng.choice(dimx * dimy, 9, replace=False)
            goals = np.stack([cells % dimx, cells // dimx], 1)
            got = capi.bfs_fields(dimx, dimy, obst, goals)
            want = orc.bfs_fields(dimx, dimy, obst, goals)
            assert np.array_equal(got, want)
    finally:
        for k in env:
            os.environ.pop(k, None)


@pytest.mark.parametrize("dimx,dimy,density", [
    (33, 32, 0.2), (100, 37, 0.35), (257, 300, 0.25), (1000, 70, 0.2), (70, 1100, 0.2),
    (1024, 130, 0.2), (640, 200, 0.45), (96, 67, -1.0)])
def test_detour_sweep_kernel(capi, orc, dimx, dimy, density):
    """MRP_BFS_SWEEP=1: the detour-level sweep (bfs_sweep.cu), opt-in because it is slower than the
    queue kernel; goals that need more detour levels than it holds (the serpentine maze, the 45 %
    map) are handed to the queue kernel.  Same bytes as the oracle either way."""
    rng = np.random.default_rng(dimx * 7 + dimy)
    if density < 0:
        obst = []
        for y in range(1, dimy, 2):
            gap = dimx - 1 if (y // 2) % 2 == 0 else 0
            obst += [[x, y] for x in range(dimx) if x != gap]
        obst = np.asarray(obst, np.int32)
    else:
        obst = _rand_map(rng, dimx, dimy, density)
    cells = rng.choice(dimx * dimy, 10, replace=False)
    goals = np.stack([cells % dimx, cells // dimx], 1)  # may sit on obstacles
    goals[0] = (0, 0)
    goals[1] = (dimx - 1, dimy - 1)
    got = _with_env({"MRP_BFS_SWEEP": "1"}, lambda: capi.bfs_fields(dimx, dimy, obst, goals))
    want = orc.bfs_fields(dimx, dimy, obst, goals)
    assert np.array_equal(got, want)


def test_empty_and_errors(capi):
    assert capi.bfs_fields(8, 8, [], np.zeros((0, 2))).shape == (0, 64)
    with pytest.raises(capi.MrpError):
        capi.bfs_fields(8, 8, [], [[8, 0]])
    with pytest.raises(capi.MrpError):
        capi.bfs_fields(0, 8, [], [[0, 0]])
    # obstacles outside the map are ignored like in the reference
    f = capi.bfs_fields(3, 1, [[5, 5], [-1, 0]], [[0, 0]])
    assert list(f[0]) == [0, 1, 2]


def test_maze_deep(capi, orc):
    # serpentine corridor: BFS depth ~ cells/2, exercises long wavefronts
    dimx, dimy = 96, 67
    obst = []
    for y in range(1, dimy, 2):
        gap = dimx - 1 if (y // 2) % 2 == 0 else 0
        obst += [[x, y] for x in range(dimx) if x != gap]
    goals = [[0, 0], [dimx - 1, dimy - 1], [40, 20]]
    got = capi.bfs_fields(dimx, dimy, obst, goals)
    want = orc.bfs_fields(dimx, dimy, obst, goals)
    assert np.array_equal(got, want)
    assert got[0][got[0] != capi.INF].max() > 2000


def _with_env(env, fn):
    try:
        os.environ.update(env)
        return fn()
    finally:
        for k in env:
            os.environ.pop(k, None)


@pytest.mark.parametrize("fmt", ["8", "16"])
@pytest.mark.parametrize("batch", ["1", "3", "0"])
def test_packed_transfer(capi, orc, batch, fmt):
    """Fields sent as detour bytes or uint16 and expanded on the host
    (mrp_bfs_fields for large results) equal the int32 transfer and the oracle;
    many small batches cycle through the three staging slots."""
    rng = np.random.default_rng(5)
    for dimx, dimy, density, ng in [(300, 200, 0.2, 40), (33, 7, 0.1, 11), (1, 1, 0.0, 1),
                                    (1024, 40, 0.2, 7), (48, 31, 0.3, 9)]:
        obst = _rand_map(rng, dimx, dimy, density)
        cells = rng.choice(dimx * dimy, ng, replace=False)
        goals = np.stack([cells % dimx, cells // dimx], 1)
        want = orc.bfs_fields(dimx, dimy, obst, goals)
        env = {"MRP_BFS_PACK": "1", "MRP_WIDEN_THREADS": "3", "MRP_BFS_FMT": fmt}
        if batch != "0":
            env["MRP_BFS_BATCH"] = batch
        # an output buffer that is only 4-byte aligned
        raw = np.empty(ng * dimx * dimy + 1, np.int32)
        out = raw[1:].reshape(ng, dimx * dimy)
        got = _with_env(env, lambda: capi.bfs_fields(dimx, dimy, obst, goals, out=out))
        assert np.array_equal(got, want)
        # random maps of this density have short detours: nothing is sent twice
        assert capi.bfs_d2h_bytes() == ng * dimx * dimy * (1 if fmt == "8" else 2)
        plain = _with_env({"MRP_BFS_PACK": "0"}, lambda: capi.bfs_fields(dimx, dimy, obst, goals))
        assert np.array_equal(plain, want)
        assert capi.bfs_d2h_bytes() == ng * dimx * dimy * 4


def _serpentine(dimx, dimy):
    obst = []
    for y in range(1, dimy, 2):
        gap = dimx - 1 if (y // 2) % 2 == 0 else 0
        obst += [[x, y] for x in range(dimx) if x != gap]
    return obst


def test_packed_transfer_overflow(capi, orc):
    """A batch with a detour of 510 steps or more cannot travel as bytes and is
    sent again as uint16; one with a finite distance >= 65535 is sent again as
    int32; the other batches stay in the narrow format."""
    dimx, dimy = 520, 261
    obst = _serpentine(dimx, dimy)
    goals = [[0, 0], [260, 130], [0, 0], [dimx - 1, dimy - 1], [3, 128]]
    want = orc.bfs_fields(dimx, dimy, obst, goals)
    assert want[0][want[0] != capi.INF].max() > 65535
    assert want[1][want[1] != capi.INF].max() < 65535
    for fmt in ("8", "16"):
        for batch in ("1", "2", "5"):
            got = _with_env({"MRP_BFS_PACK": "1", "MRP_BFS_BATCH": batch, "MRP_BFS_FMT": fmt},
                            lambda: capi.bfs_fields(dimx, dimy, obst, goals))
            assert np.array_equal(got, want), (fmt, batch)
    # a small serpentine: detours beyond a byte, distances within uint16; the
    # open map next to it in the same call stays in bytes when it has its own batch
    dimx, dimy = 64, 33
    obst = _serpentine(dimx, dimy)
    goals = [[0, 0], [63, 32], [5, 16]]
    want = orc.bfs_fields(dimx, dimy, obst, goals)
    fin = want[0] != capi.INF
    assert 510 < want[0][fin].max() < 65535
    cells = dimx * dimy
    for batch, nbytes in (("1", None), ("3", 3 * cells * (1 + 2))):
        got = _with_env({"MRP_BFS_PACK": "1", "MRP_BFS_BATCH": batch},
                        lambda: capi.bfs_fields(dimx, dimy, obst, goals))
        assert np.array_equal(got, want), batch
        if nbytes is not None:
            assert capi.bfs_d2h_bytes() == nbytes


def test_c5_map_sample(capi, orc):
    from libmultirobotplanning_b200 import instances
    inst = instances.synthetic_c5()
    assert len(inst.obstacles) == 209448
    goals = inst.goals[:6]
    got = capi.bfs_fields(1024, 1024, inst.obstacles, goals)
    want = orc.bfs_fields(1024, 1024, inst.obstacles, goals)
    assert np.array_equal(got, want)


def test_c5_properties_many_goals(capi):
    """Full-size properties that need no oracle: 0 exactly at the goal, MRP_INF
    exactly on obstacles/other components, neighbouring free cells differ by
    exactly 1 (bipartite grid), symmetry d(a,b) == d(b,a)."""
    from libmultirobotplanning_b200 import instances
    inst = instances.synthetic_c5()
    G = 300
    goals = inst.goals[:G]
    f = capi.bfs_fields(1024, 1024, inst.obstacles, goals).reshape(G, 1024, 1024)
    free = np.ones((1024, 1024), bool)
    free[inst.obstacles[:, 1], inst.obstacles[:, 0]] = False
    reach = f[0] != capi.INF
    assert reach.sum() == 837416
    for g in range(G):
        fg = f[g]
        assert fg[goals[g, 1], goals[g, 0]] == 0
        assert ((fg != capi.INF) == reach).all()
        assert (fg == 0).sum() == 1
    for g in range(0, G, 37):
        fg = f[g].astype(np.int64)
        both = reach[:, 1:] & reach[:, :-1]
        assert (np.abs(fg[:, 1:] - fg[:, :-1])[both] == 1).all()
        both = reach[1:, :] & reach[:-1, :]
        assert (np.abs(fg[1:, :] - fg[:-1, :])[both] == 1).all()
    for a in range(0, 40, 7):
        for b in range(1, 40, 11):
            assert f[a][goals[b, 1], goals[b, 0]] == f[b][goals[a, 1], goals[a, 0]]


def test_device_entry_point(capi, orc):
    import torch
    dimx, dimy = 200, 150
    rng = np.random.default_rng(5)
    obst = _rand_map(rng, dimx, dimy, 0.2)
    cells = rng.choice(dimx * dimy, 40, replace=False).astype(np.int32)
    m = capi.Map(dimx, dimy, obst)
    d_goals = torch.from_numpy(cells).cuda()
    d_out = torch.empty((40, dimx * dimy), dtype=torch.int32, device="cuda")
    ws = torch.empty(max(m.workspace_bytes(40), 256), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream()
    m.bfs_fields_dev(d_goals.data_ptr(), 40, d_out.data_ptr(), ws.data_ptr(), s.cuda_stream)
    s.synchronize()
    goals = np.stack([cells % dimx, cells // dimx], 1)
    assert np.array_equal(d_out.cpu().numpy(), orc.bfs_fields(dimx, dimy, obst, goals))
    m.close()


def test_packed_result_mode(capi, orc):
    """mrp_bfs_fields_packed: one detour byte per cell, decoded with the accessor formula of
    mrp_packed_value (2*b + |x-gx| + |y-gy|, 255 = INF) == the oracle's int fields; a maze whose
    detours do not fit a byte is reported per goal."""
    rng = np.random.default_rng(11)
    for dimx, dimy, density, batch in [(300, 200, 0.2, None), (1024, 96, 0.2, "3"), (33, 45, 0.3, "2")]:
        obst = _rand_map(rng, dimx, dimy, density)
        cells = rng.choice(dimx * dimy, 7, replace=False)
        goals = np.stack([cells % dimx, cells // dimx], 1)
        env = {"MRP_BFS_BATCH": batch} if batch else {}
        packed, ovf = _with_env(env, lambda: capi.bfs_fields_packed(dimx, dimy, obst, goals))
        assert ovf.sum() == 0
        want = orc.bfs_fields(dimx, dimy, obst, goals)
        assert np.array_equal(capi.unpack_field(packed, dimx, dimy, goals), want)
    # serpentine corridor: detours far beyond 2 * 254
    dimx, dimy = 96, 67
    obst = []
    for y in range(1, dimy, 2):
        gap = dimx - 1 if (y // 2) % 2 == 0 else 0
        obst += [[x, y] for x in range(dimx) if x != gap]
    goals = [[0, 0], [dimx - 1, dimy - 1], [40, 20]]
    packed, ovf = capi.bfs_fields_packed(dimx, dimy, obst, goals)
    assert ovf.all()


@pytest.mark.parametrize("batch", ["0", "3"])
def test_compact_result_mode(capi, orc, batch):
    """mrp_bfs_fields_compact: detour bytes of the free cells only, positions from
    mrp_free_cell_index; read back with the numpy mirror of mrp_compact_value it must equal
    the oracle's int fields on odd map shapes (cell counts that are no multiple of 32,
    single-cell map, no obstacles, goals on obstacles); the D2H bytes are n_free per field;
    a serpentine whose detours do not fit a byte is reported as overflowed."""
    rng = np.random.default_rng(15)
    for dimx, dimy, density, ng in [(300, 200, 0.2, 40), (33, 7, 0.1, 11), (1, 1, 0.0, 1),
                                    (1024, 40, 0.2, 7), (48, 31, 0.0, 9), (64, 64, 0.45, 20)]:
        obst = _rand_map(rng, dimx, dimy, density)
        cells = rng.choice(dimx * dimy, ng, replace=False)
        goals = np.stack([cells % dimx, cells // dimx], 1)
        want = orc.bfs_fields(dimx, dimy, obst, goals)
        bits, prefix, n_free = capi.free_cell_index(dimx, dimy, obst)
        env = {} if batch == "0" else {"MRP_BFS_BATCH": batch}
        got, ovf = _with_env(env, lambda: capi.bfs_fields_compact(dimx, dimy, obst, goals, n_free))
        assert got.shape == (ng, n_free) and not ovf.any()
        assert capi.bfs_d2h_bytes() == ng * n_free
        assert np.array_equal(capi.unpack_compact(got, bits, dimx, dimy, goals), want)
    dimx, dimy = 64, 33
    obst = _serpentine(dimx, dimy)
    goals = [[0, 0], [63, 32], [5, 16]]
    bits, prefix, n_free = capi.free_cell_index(dimx, dimy, obst)
    got, ovf = _with_env({"MRP_BFS_BATCH": "1"}, lambda: capi.bfs_fields_compact(dimx, dimy, obst, goals, n_free))
    assert ovf[0] == 1
    want = orc.bfs_fields(dimx, dimy, obst, goals)
    ok = [k for k in range(3) if not ovf[k]]
    assert np.array_equal(capi.unpack_compact(got[ok], bits, dimx, dimy, np.asarray(goals)[ok]), want[ok])


def test_fields_equal_reference_class(capi, set8, set32):
    """The CUDA fields against the reference's OWN ShortestPathHeuristic class
    (example/shortest_path_heuristic.hpp:12-62 compiled unmodified; CRCs committed by
    tests/golden/make_sph_golden.py): benchmark maps and seeded odd shapes, bit for bit."""
    import json
    import os
    import sys
    gdir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    if gdir not in sys.path:
        sys.path.insert(0, gdir)
    import make_sph_golden as S
    g = json.load(open(os.path.join(gdir, "sph_fields_golden.json")))
    for name, (dx, dy, obst, goals) in S.cases(set8, set32).items():
        f = capi.bfs_fields(dx, dy, obst, goals)
        assert S.crc(f) == g[name]["crc32"], name
