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
    return sorted(set(re.findall(r"\b(mrp_[a-z0-9_]+)\s*\(", text)))


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


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "libmultirobotplanning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in text.lower().replace("# oracle", ""), f
