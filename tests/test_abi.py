"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every
symbol include/mrp_b200.h declares, and fails loudly (no CPU fallback) when
there is no CUDA device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mrp_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    # header-only helpers (static inline) are not exported symbols
    inline = set(re.findall(r"static\s+inline\s+[^;{(]*?\b(mrp_[a-z0-9_]+)\s*\(", text))
    return sorted(set(re.findall(r"\b(mrp_[a-z0-9_]+)\s*\(", text)) - inline)


def test_header_symbols_exported():
    from libmultirobotplanning_b200 import capi
    lib = capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 18
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(capi.EXPORTS) == declared


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from libmultirobotplanning_b200 import capi
    with pytest.raises(capi.MrpError) as e:
        capi.bfs_fields(4, 4, [], [[0, 0]])
    assert e.value.code == -1 and "no CPU fallback" in str(e.value)
    with pytest.raises(capi.MrpError):
        capi.count_conflicts([[0, 1], [1, 0]], [2, 2])


def test_widen_u16_host_half_of_packed_transfer():
    """mrp_widen_u16 (no device needed): 0xFFFF -> MRP_INF, everything else
    unchanged, for ragged sizes, unaligned buffers and any thread count."""
    import numpy as np
    from libmultirobotplanning_b200 import capi
    rng = np.random.default_rng(3)
    for n in (0, 1, 7, 15, 16, 17, 1000, 65536, 200003, 1 << 21):
        for threads in (1, 5):
            for off in (0, 1, 3):
                src = rng.integers(0, 65536, n + off, dtype=np.uint16)[off:]
                if n:
                    src[rng.integers(0, n, max(1, n // 9))] = 0xFFFF
                raw = np.full(n + off + 8, -7, np.int32)
                dst = raw[off:off + n]
                capi.widen_u16(src, threads, out=dst)
                want = np.where(src == 0xFFFF, capi.INF, src.astype(np.int32))
                assert np.array_equal(dst, want), (n, threads, off)
                assert (raw[:off] == -7).all() and (raw[off + n:] == -7).all()


def test_widen_u8_host_half_of_packed_transfer():
    """mrp_widen_u8 (no device needed): distance = 2*byte + Manhattan distance
    to the field's goal, 255 -> MRP_INF, for ragged widths, unaligned buffers
    and any thread count; round trip from oracle fields."""
    import numpy as np
    from libmultirobotplanning_b200 import capi
    from oracle import orc
    rng = np.random.default_rng(4)
    for dimx, dimy, ng in [(1, 1, 1), (7, 5, 3), (16, 16, 4), (33, 9, 5), (300, 70, 6), (1024, 20, 3)]:
        cells = dimx * dimy
        free = rng.random(cells) > 0.2
        obst = [[c % dimx, c // dimx] for c in np.flatnonzero(~free)]
        goal_cell = rng.choice(cells, ng, replace=False).astype(np.int32)
        goals = np.stack([goal_cell % dimx, goal_cell // dimx], 1)
        want = orc.bfs_fields(dimx, dimy, obst, goals)
        xs, ys = np.arange(cells) % dimx, np.arange(cells) // dimx
        man = np.abs(xs[None] - goals[:, :1]) + np.abs(ys[None] - goals[:, 1:])
        fin = want != capi.INF
        det = np.where(fin, want - man, 0)
        assert (det >= 0).all() and (det % 2 == 0).all()
        assert det.max() // 2 < 255
        src = np.where(fin, det // 2, 255).astype(np.uint8)
        for threads in (1, 5):
            for off in (0, 1, 3):
                raw = np.full(ng * cells + off + 8, -7, np.int32)
                dst = raw[off:off + ng * cells].reshape(ng, cells)
                capi.widen_u8(src, dimx, dimy, goal_cell, threads, out=dst)
                assert np.array_equal(dst, want), (dimx, dimy, threads, off)
                assert (raw[:off] == -7).all() and (raw[off + ng * cells:] == -7).all()
    with pytest.raises(capi.MrpError):
        capi.widen_u8(np.zeros((1, 4), np.uint8), 2, 2, [4])


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "libmultirobotplanning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("# oracle", ""), f


def test_free_cell_index_without_a_gpu():
    """mrp_free_cell_index (the index of the compact field format) is host code: free mask,
    prefix sums and the number of free cells against numpy, on a map whose cell count is not
    a multiple of 32, obstacles outside the map ignored."""
    import numpy as np
    from libmultirobotplanning_b200 import capi
    rng = np.random.default_rng(4)
    dimx, dimy = 37, 21
    blocked = rng.random((dimy, dimx)) < 0.3
    ys, xs = np.nonzero(blocked)
    obst = np.concatenate([np.stack([xs, ys], 1), [[-1, 3], [dimx, 0], [5, dimy]]]).astype(np.int32)
    bits, prefix, n = capi.free_cell_index(dimx, dimy, obst)
    free = (~blocked).ravel()
    assert n == free.sum() == prefix[-1]
    got = np.unpackbits(bits.view(np.uint8), bitorder="little")
    assert (got[:dimx * dimy].astype(bool) == free).all() and not got[dimx * dimy:].any()
    cs = np.concatenate([[0], np.cumsum(free)])
    assert all(prefix[w] == cs[min(32 * w, dimx * dimy)] for w in range(len(prefix)))


def test_bitmap_row_division_constants():
    """The queue BFS kernel turns a bitmap word index into a row with a 32-bit multiply-shift
    (bfs_queue.cu: umulhi(word, magic) >> shift).  For every map width: the row stride is odd, at
    least three words and covers the columns plus the border, and the constants divide exactly for
    every word index the kernel can form (< 2^27).  (Maps at most 30 columns wide once got a stride of
    one word, for which no 32-bit constant exists: narrow maps of more than one tile came out wrong.)
    Host code only: no device needed."""
    import ctypes as C
    import numpy as np
    from libmultirobotplanning_b200 import _capi
    lib = _capi.lib()
    rng = np.random.default_rng(0)
    widths = list(range(1, 200)) + [255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 2047, 2048, 2049, 4096,
                                    30000, 65533]
    for dimx in widths:
        rw, mg, sh = C.c_int32(), C.c_uint32(), C.c_int32()
        assert lib.mrp_bitmap_row_division(dimx, C.byref(rw), C.byref(mg), C.byref(sh)) == 0
        W = rw.value
        assert W % 2 == 1 and W >= 3 and 32 * W >= dimx + 2, dimx
        k = np.concatenate([np.arange(0, 4000, dtype=np.uint64),
                            rng.integers(0, 1 << 27, 8000).astype(np.uint64),
                            np.array([(1 << 27) - 1], np.uint64)])
        m = (k // np.uint64(W)) * np.uint64(W)  # both sides of every row boundary
        w = np.unique(np.concatenate([k, m, np.maximum(m, np.uint64(1)) - np.uint64(1), m + np.uint64(W - 1)]))
        w = w[w < (1 << 27)]
        got = ((w * np.uint64(mg.value)) >> np.uint64(32)) >> np.uint64(sh.value)
        assert np.array_equal(got, w // np.uint64(W)), dimx
    assert lib.mrp_bitmap_row_division(0, C.byref(rw), C.byref(mg), C.byref(sh)) < 0
