"""sgc_b200 — B200-native hot paths of StaggeredGridCalculus.jl (matrix-free operators and
sparse assembly) behind the reference's own API.

The directory is named `staggeredgridcalculus.jl_b200/` (not an importable identifier); the
top-level module `sgc_b200.py` loads it under the name `sgc_b200`.  Importing requires the
in-tree CUDA library `csrc/libsgc_b200.so` — there is no CPU fallback.
"""
from . import _lib
from ._lib import SgcError, InexactError, SGC_F64, SGC_C128, LIB_PATH
from .device import DeviceArray
from .matrixfree import (Operator, apply_partial, apply_mhat, apply_curl, apply_curlcurl, apply_divg, apply_grad,
                         apply_mean)
from .matrix import (DeviceCSC, create_partial, create_mhat, create_mean, create_curl, create_divg, create_grad,
                     create_picmp)
from . import gridgen


def launch_count() -> int:
    """Kernels launched by libsgc_b200 in this process so far."""
    return int(_lib.lib.sgc_launch_count())
