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
                                        (513, 64, 5000)])
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
