"""ctypes binding of oracle/liboracle.so + numpy helpers.  TEST INFRASTRUCTURE ONLY: imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs — never by the product."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        i, sz, u64, vp = C.c_int, C.c_size_t, C.c_uint64, C.c_void_p
        for s, ct in (("d", C.c_double), ("s", C.c_float)):
            getattr(lib, f"oracle_{s}gemm").argtypes = [i, i, i, i, i, ct, vp, i, vp, i, ct, vp, i]
            getattr(lib, f"oracle_{s}gemm").restype = None
            getattr(lib, f"oracle_{s}gemm_exact").argtypes = [i, i, i, i, i, ct, vp, i, vp, i, ct, vp, i, vp, vp]
            getattr(lib, f"oracle_{s}gemm_exact").restype = None
            getattr(lib, f"oracle_{s}geam").argtypes = [i, i, i, i, ct, vp, i, ct, vp, i, vp, i]
            getattr(lib, f"oracle_{s}geam").restype = None
            getattr(lib, f"oracle_{s}gemm_plan").argtypes = [C.c_char_p, i, i, i, i, i, ct, vp, i, vp, i, ct, vp, i, vp, sz]
            getattr(lib, f"oracle_{s}gemm_plan").restype = i
            getattr(lib, f"oracle_{s}syrk").argtypes = [i, i, i, i, ct, vp, i, ct, vp, i]
            getattr(lib, f"oracle_{s}syrk").restype = None
            getattr(lib, f"oracle_{s}trsm").argtypes = [i, i, i, i, i, i, ct, vp, i, vp, i]
            getattr(lib, f"oracle_{s}trsm").restype = None
            getattr(lib, f"oracle_{s}syrk_plan").argtypes = [C.c_char_p, i, i, i, i, ct, vp, i, ct, vp, i, vp, sz]
            getattr(lib, f"oracle_{s}syrk_plan").restype = i
            getattr(lib, f"oracle_{s}trsm_plan").argtypes = [C.c_char_p, i, i, i, i, i, i, ct, vp, i, vp, i, vp, sz]
            getattr(lib, f"oracle_{s}trsm_plan").restype = i
            getattr(lib, f"oracle_fill_uniform_{s}").argtypes = [vp, sz, u64, ct, ct]
            getattr(lib, f"oracle_fill_uniform_{s}").restype = None
        lib.oracle_gemm_plan_workspace.argtypes = [C.c_char_p, i, i, i, i, i, sz]
        lib.oracle_gemm_plan_workspace.restype = sz
        lib.oracle_syrk_plan_workspace.argtypes = [C.c_char_p, i, i, i, sz]
        lib.oracle_syrk_plan_workspace.restype = sz
        lib.oracle_trsm_plan_workspace.argtypes = [C.c_char_p, i, i, i, sz]
        lib.oracle_trsm_plan_workspace.restype = sz

    @staticmethod
    def _s(dtype):
        return "d" if np.dtype(dtype) == np.float64 else "s"

    def fill(self, n, seed, dtype=np.float64, lo=-1.0, hi=1.0):
        x = np.empty(n, dtype=dtype)
        getattr(self.lib, f"oracle_fill_uniform_{self._s(dtype)}")(x.ctypes.data, n, seed, lo, hi)
        return x

    def gemm(self, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc):
        getattr(self.lib, f"oracle_{self._s(Cm.dtype)}gemm")(ta, tb, m, n, k, alpha, A.ctypes.data, lda, B.ctypes.data, ldb, beta, Cm.ctypes.data, ldc)
        return Cm

    def gemm_exact(self, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C0, ldc):
        exact = np.empty(m * n, dtype=np.float64)
        absab = np.empty(m * n, dtype=np.float64)
        c0 = C0.ctypes.data if C0 is not None else None
        getattr(self.lib, f"oracle_{self._s(A.dtype)}gemm_exact")(ta, tb, m, n, k, alpha, A.ctypes.data, lda, B.ctypes.data, ldb, beta, c0, ldc, exact.ctypes.data, absab.ctypes.data)
        return exact.reshape(n, m).T, absab.reshape(n, m).T

    def geam(self, ta, tb, m, n, alpha, A, lda, beta, B, ldb, Cm, ldc):
        a = A.ctypes.data if A is not None else None
        b = B.ctypes.data if B is not None else None
        getattr(self.lib, f"oracle_{self._s(Cm.dtype)}geam")(ta, tb, m, n, alpha, a, lda, beta, b, ldb, Cm.ctypes.data, ldc)
        return Cm

    def plan_workspace(self, plan, ta, tb, m, n, k, elem):
        return self.lib.oracle_gemm_plan_workspace(plan.encode(), ta, tb, m, n, k, elem)

    def gemm_plan(self, plan, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc, ws=None):
        need = self.plan_workspace(plan, ta, tb, m, n, k, Cm.dtype.itemsize)
        if ws is None:
            ws = np.full(need // Cm.dtype.itemsize + 1, np.nan, dtype=Cm.dtype)
            nbytes = need
        else:
            nbytes = ws.nbytes
        rc = getattr(self.lib, f"oracle_{self._s(Cm.dtype)}gemm_plan")(plan.encode(), ta, tb, m, n, k, alpha, A.ctypes.data, lda, B.ctypes.data, ldb, beta, Cm.ctypes.data, ldc, ws.ctypes.data, nbytes)
        return rc, ws


    # ---- SYRK / TRSM (flags are 0/1: lower, trans, left, unit) ------------------------------------
    def syrk(self, lower, trans, n, k, alpha, A, lda, beta, Cm, ldc):
        getattr(self.lib, f"oracle_{self._s(Cm.dtype)}syrk")(lower, trans, n, k, alpha, A.ctypes.data, lda, beta, Cm.ctypes.data, ldc)
        return Cm

    def trsm(self, left, lower, trans, unit, m, n, alpha, A, lda, B, ldb):
        getattr(self.lib, f"oracle_{self._s(B.dtype)}trsm")(left, lower, trans, unit, m, n, alpha, A.ctypes.data, lda, B.ctypes.data, ldb)
        return B

    def syrk_plan_workspace(self, plan, trans, n, k, elem):
        return self.lib.oracle_syrk_plan_workspace(plan.encode(), trans, n, k, elem)

    def trsm_plan_workspace(self, plan, left, m, n, elem):
        return self.lib.oracle_trsm_plan_workspace(plan.encode(), left, m, n, elem)

    def syrk_plan(self, plan, lower, trans, n, k, alpha, A, lda, beta, Cm, ldc):
        need = self.syrk_plan_workspace(plan, trans, n, k, Cm.dtype.itemsize)
        ws = np.full(need // Cm.dtype.itemsize + 1, np.nan, dtype=Cm.dtype)
        rc = getattr(self.lib, f"oracle_{self._s(Cm.dtype)}syrk_plan")(plan.encode(), lower, trans, n, k, alpha, A.ctypes.data, lda, beta,
                                                                     Cm.ctypes.data, ldc, ws.ctypes.data, need)
        return rc, ws

    def trsm_plan(self, plan, left, lower, trans, unit, m, n, alpha, A, lda, B, ldb):
        need = self.trsm_plan_workspace(plan, left, m, n, B.dtype.itemsize)
        ws = np.full(need // B.dtype.itemsize + 1, np.nan, dtype=B.dtype)
        rc = getattr(self.lib, f"oracle_{self._s(B.dtype)}trsm_plan")(plan.encode(), left, lower, trans, unit, m, n, alpha, A.ctypes.data, lda,
                                                                    B.ctypes.data, ldb, ws.ctypes.data, need)
        return rc, ws


_cached = None


def load():
    global _cached
    if _cached is None:
        so = os.path.join(ORACLE_DIR, "liboracle.so")
        if not os.path.exists(so):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"])
        _cached = Oracle(C.CDLL(so))
    return _cached


# ---- column-major helpers -------------------------------------------------------------------------
def colmajor(rows, cols, ld, seed, dtype=np.float64, orc=None):
    """1-D buffer of ld*cols elements filled by the shared counter-based generator."""
    return (orc or load()).fill(ld * cols, seed, dtype)


def view(buf, rows, cols, ld):
    """rows x cols numpy view (Fortran order) of a column-major buffer with leading dimension ld."""
    return np.lib.stride_tricks.as_strided(buf, shape=(rows, cols), strides=(buf.itemsize, ld * buf.itemsize))


BOOL_PLANS = ["FF", "FT", "TF", "TT"]  # SYRK_Options / TRSM_Options
GEMM_PLANS = [a + b + c for a in "NT" for b in "NT" for c in "NT"]
GEMM_PAD_PLANS = [a + b + c + d + e + f for a in "NT" for b in "NT" for c in "NT" for d in "NP" for e in "NP" for f in "NP"]
