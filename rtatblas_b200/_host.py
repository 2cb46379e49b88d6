"""ctypes binding of include/rtat_b200_host.h: the C++ planner/executor (rtat.h surface) for Python."""
import ctypes as C
import json
import os

from ._capi import RtatError, _ptr, Handle

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def host_lib_path():
    return os.path.join(_HERE, "lib", "librtat_host.so")


def host_lib():
    global _lib
    if _lib is None:
        p = host_lib_path()
        if not os.path.exists(p):
            raise RtatError(f"{p} is missing: build it with `make` (or __graft_entry__.build())")
        L = C.CDLL(p)
        vp, i, sz, ll, cp = C.c_void_p, C.c_int, C.c_size_t, C.c_longlong, C.c_char_p
        L.rtat_host_planner_create.restype = vp
        L.rtat_host_planner_create.argtypes = [cp, cp]
        L.rtat_host_planner_destroy.argtypes = [vp]
        L.rtat_host_planner_gemm.argtypes = [vp, vp, i, i, i, i, i, vp, vp, i, vp, i, vp, vp, i, vp, sz, cp, cp]
        L.rtat_host_planner_workspace.restype = ll
        L.rtat_host_planner_workspace.argtypes = [vp, i, i, i, i, i, i, i, i, cp]
        L.rtat_host_planner_syrk.argtypes = [vp, vp, i, i, i, i, vp, vp, i, vp, vp, i, vp, sz, cp, cp]
        L.rtat_host_planner_syrk_workspace.restype = ll
        L.rtat_host_planner_syrk_workspace.argtypes = [vp, i, i, i, i, i, i, cp]
        L.rtat_host_planner_trsm.argtypes = [vp, vp, i, i, i, i, i, i, vp, vp, i, vp, i, vp, sz, cp, cp]
        L.rtat_host_planner_trsm_workspace.restype = ll
        L.rtat_host_planner_trsm_workspace.argtypes = [vp, i, i, i, i, i, i, i, i, cp]
        for n in ("statistics", "save_plans"):
            getattr(L, f"rtat_host_planner_{n}").restype = vp
            getattr(L, f"rtat_host_planner_{n}").argtypes = [vp]
        L.rtat_host_planner_load_plans.restype = ll
        L.rtat_host_planner_load_plans.argtypes = [vp, cp]
        L.rtat_host_planner_set_tests_until_converge.argtypes = [vp, i]
        L.rtat_host_planner_save_plans_file.argtypes = [vp, cp]
        L.rtat_host_planner_load_plans_file.restype = ll
        L.rtat_host_planner_load_plans_file.argtypes = [vp, cp]
        L.rtat_host_planner_set_default_preference.argtypes = [vp, C.c_double]
        L.rtat_host_planner_statistics_rates.restype = vp
        L.rtat_host_planner_statistics_rates.argtypes = [vp]
        L.rtat_host_planner_num_plans.argtypes = [vp]
        L.rtat_host_planner_plan_name.restype = vp
        L.rtat_host_planner_plan_name.argtypes = [vp, i]
        L.rtat_host_free.argtypes = [vp]
        L.rtat_host_last_error.restype = cp
        _lib = L
    return _lib


def _take(ptr):
    if not ptr:
        raise RtatError(host_lib().rtat_host_last_error().decode())
    text = C.string_at(ptr).decode()
    host_lib().rtat_host_free(ptr)
    return text


class Planner:
    """Planning_System<GEMM_Executor[_Pad]<T>> (reference src/planning/planning_system.h:38-131)."""

    def __init__(self, method="gemm", data_type="double"):
        self.double = data_type == "double"
        self._p = host_lib().rtat_host_planner_create(method.encode(), data_type.encode())
        if not self._p:
            raise RtatError(host_lib().rtat_host_last_error().decode())

    def close(self):
        if self._p:
            host_lib().rtat_host_planner_destroy(self._p)
            self._p = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def plans(self):
        n = host_lib().rtat_host_planner_num_plans(self._p)
        return [_take(host_lib().rtat_host_planner_plan_name(self._p, i)) for i in range(n)]

    def workspace(self, transa, transb, m, n, k, plan, lda=None, ldb=None, ldc=None):
        lda = lda or (k if transa else m)
        ldb = ldb or (n if transb else k)
        ldc = ldc or m
        r = host_lib().rtat_host_planner_workspace(self._p, transa, transb, m, n, k, lda, ldb, ldc, plan.encode())
        if r < 0:
            raise RtatError(host_lib().rtat_host_last_error().decode())
        return r

    def gemm(self, handle: Handle, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc, workspace=None,
             workspace_bytes=0, plan=None):
        """create_plan (unless `plan` forces one) + execute on the handle's stream; returns the plan used."""
        ct = C.c_double if self.double else C.c_float
        al, be = ct(alpha), ct(beta)
        out = C.create_string_buffer(16)
        rc = host_lib().rtat_host_planner_gemm(self._p, handle.raw, transa, transb, m, n, k, C.byref(al), _ptr(A), lda,
                                               _ptr(B), ldb, C.byref(be), _ptr(Cm), ldc, _ptr(workspace), workspace_bytes,
                                               plan.encode() if plan else None, out)
        if rc != 0:
            raise RtatError("planner gemm failed: " + host_lib().rtat_host_last_error().decode())
        return out.value.decode()

    def syrk_workspace(self, uplo, trans, n, k, plan, lda=None, ldc=None):
        r = host_lib().rtat_host_planner_syrk_workspace(self._p, uplo, trans, n, k, lda or (k if trans else n), ldc or n, plan.encode())
        if r < 0:
            raise RtatError(host_lib().rtat_host_last_error().decode())
        return r

    def syrk(self, handle: Handle, uplo, trans, n, k, alpha, A, lda, beta, Cm, ldc, workspace=None, workspace_bytes=0, plan=None):
        """SYRK through Planning_System<SYRK_Executor>; returns the plan used ("FF", "FT", "TF" or "TT")."""
        ct = C.c_double if self.double else C.c_float
        al, be = ct(alpha), ct(beta)
        out = C.create_string_buffer(16)
        rc = host_lib().rtat_host_planner_syrk(self._p, handle.raw, uplo, trans, n, k, C.byref(al), _ptr(A), lda, C.byref(be), _ptr(Cm), ldc,
                                               _ptr(workspace), workspace_bytes, plan.encode() if plan else None, out)
        if rc != 0:
            raise RtatError("planner syrk failed: " + host_lib().rtat_host_last_error().decode())
        return out.value.decode()

    def trsm_workspace(self, side, uplo, trans, diag, m, n, plan, lda=None, ldb=None):
        r = host_lib().rtat_host_planner_trsm_workspace(self._p, side, uplo, trans, diag, m, n, lda or (n if side else m), ldb or m,
                                                        plan.encode())
        if r < 0:
            raise RtatError(host_lib().rtat_host_last_error().decode())
        return r

    def trsm(self, handle: Handle, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, workspace=None, workspace_bytes=0, plan=None):
        """TRSM through Planning_System<TRSM_Executor>; returns the plan used."""
        al = (C.c_double if self.double else C.c_float)(alpha)
        out = C.create_string_buffer(16)
        rc = host_lib().rtat_host_planner_trsm(self._p, handle.raw, side, uplo, trans, diag, m, n, C.byref(al), _ptr(A), lda, _ptr(B), ldb,
                                               _ptr(workspace), workspace_bytes, plan.encode() if plan else None, out)
        if rc != 0:
            raise RtatError("planner trsm failed: " + host_lib().rtat_host_last_error().decode())
        return out.value.decode()

    def statistics(self):
        return json.loads(_take(host_lib().rtat_host_planner_statistics(self._p)))

    def save_plans(self):
        return json.loads(_take(host_lib().rtat_host_planner_save_plans(self._p)))

    def load_plans(self, plans):
        r = host_lib().rtat_host_planner_load_plans(self._p, json.dumps(plans).encode())
        if r < 0:
            raise RtatError(host_lib().rtat_host_last_error().decode())
        return r

    def set_tests_until_converge(self, n):
        host_lib().rtat_host_planner_set_tests_until_converge(self._p, n)

    def set_default_preference(self, margin):
        """keep the default (fused) plan unless another one is faster by more than `margin` (0 = the reference's strict minimum)"""
        host_lib().rtat_host_planner_set_default_preference(self._p, float(margin))

    def save_plans_file(self, path):
        if host_lib().rtat_host_planner_save_plans_file(self._p, str(path).encode()) != 0:
            raise RtatError(host_lib().rtat_host_last_error().decode())

    def load_plans_file(self, path):
        r = host_lib().rtat_host_planner_load_plans_file(self._p, str(path).encode())
        if r < 0:
            raise RtatError(host_lib().rtat_host_last_error().decode())
        return r

    def statistics_rates(self):
        """statistics() plus per-sample "tflops" (and "gbs" for GEMM keys) beside every plan's "times" list"""
        return json.loads(_take(host_lib().rtat_host_planner_statistics_rates(self._p)))
