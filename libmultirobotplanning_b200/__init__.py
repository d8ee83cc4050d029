"""B200-native hot path of libMultiRobotPlanning: BFS distance fields, conflict
detection/counting and batched low-level replans as hand-written sm_100a CUDA
kernels behind a C ABI (include/mrp_b200.h).  Python here is a thin ctypes
mirror used by the tests and the benchmark; the C++ host adapters and the
cbs / ecbs / cbs_ta command lines live in host/."""
from . import _capi as capi  # noqa: F401
from . import instances  # noqa: F401
from . import solver  # noqa: F401
from . import validate  # noqa: F401
