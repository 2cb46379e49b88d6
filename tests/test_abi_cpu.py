"""CPU-side checks of the drop-in boundary: the C-ABI library exists, loads, and exports every
symbol include/*.h declares.  No compute calls (there is no GPU here and no CPU fallback)."""
import ctypes
import os
import subprocess

import rtatblas_b200 as rt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    L = rt.lib()
    names = rt.exported_symbols()
    assert len(names) >= 18
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, f"declared in include/ but not exported: {missing}"
    assert L.rtat_b200_version() >= 100


def test_no_vendor_blas_or_oracle_linked_into_the_product():
    out = subprocess.check_output(["ldd", rt.lib_path()], text=True)
    for banned in ("cublas", "cublasLt", "liboracle", "openblas", "cutlass"):
        assert banned not in out, f"product library links {banned}:\n{out}"


def test_status_strings():
    L = rt.lib()
    assert L.rtat_b200_status_string(0) == b"RTAT_B200_STATUS_SUCCESS"
    assert L.rtat_b200_status_string(7) == b"RTAT_B200_STATUS_INVALID_VALUE"
    assert L.rtat_b200_status_string(15) == b"RTAT_B200_STATUS_NOT_SUPPORTED"


def test_null_handle_is_rejected_not_crashed():
    L = rt.lib()
    one = ctypes.c_double(1.0)
    st = L.rtat_b200_dgemm(None, 0, 0, 4, 4, 4, ctypes.byref(one), None, 4, None, 4, ctypes.byref(one), None, 4)
    assert st == 1  # NOT_INITIALIZED


def test_create_fails_loudly_without_a_gpu():
    import torch

    if torch.cuda.is_available():
        return  # covered by the gpu tests
    h = ctypes.c_void_p()
    st = rt.lib().rtat_b200_create(ctypes.byref(h))
    assert st != 0 and not h.value, "creating a handle without a GPU must fail (no CPU fallback)"


def test_product_sources_never_reference_the_oracle():
    bad = []
    for base, _, files in os.walk(os.path.join(ROOT, "rtatblas_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(base, f), errors="ignore").read()
                if "liboracle" in text or "oracle_binding" in text or "oracle.h" in text:
                    bad.append(os.path.join(base, f))
    assert not bad, f"product files reference the oracle: {bad}"
