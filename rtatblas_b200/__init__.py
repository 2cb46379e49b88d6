"""rtatblas_b200 — B200-native back end for rtatblas' GEMM execution path.

The product is the C-ABI library ``rtatblas_b200/lib/librtat_b200.so`` (hand-written sm_100a kernels,
``include/rtat_b200.h``) plus the C++ host layer under ``rtatblas_b200/include/rtat`` that mirrors the
reference's ``rtat.h`` surface.  This Python package is only the thin ctypes binding that tests and
``bench.py`` use; PyTorch provides device memory and streams, nothing else.

There is no CPU fallback: loading fails loudly when the CUDA library is missing, and every entry
point returns a non-zero status (raised as ``RtatError``) when no B200 is present.
"""
from ._capi import (  # noqa: F401
    RtatError,
    Handle,
    OP_N,
    OP_T,
    lib,
    lib_path,
    exported_symbols,
    host_schedule,
)
from ._host import Planner, host_lib, host_lib_path  # noqa: F401
from . import _mg as mg  # noqa: F401,E402
from ._mg import ColumnSplit  # noqa: F401,E402
