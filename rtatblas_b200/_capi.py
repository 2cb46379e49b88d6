"""ctypes binding of include/rtat_b200.h (the drop-in C ABI).  No compute happens in Python."""
import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)
OP_N, OP_T = 0, 1


class RtatError(RuntimeError):
    pass


def lib_path():
    # RTAT_B200_LIB: an experiment build of the same library (probe/build_variants.sh); never a different back end
    return os.environ.get("RTAT_B200_LIB") or os.path.join(_HERE, "lib", "librtat_b200.so")


_lib = None


def lib():
    """Load librtat_b200.so (built in-tree by `make` / __graft_entry__.build()).  Fails loudly."""
    global _lib
    if _lib is None:
        p = lib_path()
        if not os.path.exists(p):
            raise RtatError(
                f"{p} is missing: build it with `make` (or __graft_entry__.build()); there is no fallback path"
            )
        _lib = C.CDLL(p)
        _declare(_lib)
    return _lib


def host_schedule(elem_size, transa, m, n, k, lda, world=0, panels=0, lead=0):
    """The schedule the host-buffer entry points would use (rtat_b200_host_schedule; pure host arithmetic, no GPU needed):
    dict with the column blocks [(col0, cols)], how many leading ones take the K-panel phase, and the K-panels [(k0, kc)]."""
    out = (C.c_int * 128)()
    got = lib().rtat_b200_host_schedule(elem_size, transa, m, n, k, lda, world, panels, lead, out, 128)
    if got < 0:
        raise RtatError("rtat_b200_host_schedule: bad argument")
    blocks, nlead, npanels, nb = out[0], out[1], out[2], out[3]
    o = 4
    col0 = list(out[o:o + blocks]); o += blocks
    cols = list(out[o:o + blocks]); o += blocks
    k0 = list(out[o:o + npanels]); o += npanels
    kc = list(out[o:o + npanels])
    return {"blocks": list(zip(col0, cols)), "lead": nlead, "panels": list(zip(k0, kc)), "nb": nb}


def exported_symbols():
    """Names declared in include/rtat_b200*.h (used by the CPU test that checks the ABI surface)."""
    names = []
    inc = os.path.join(_ROOT, "include")
    for fn in sorted(os.listdir(inc)):
        if fn.endswith(".h"):
            text = open(os.path.join(inc, fn)).read()
            text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
            names += re.findall(r"\b(rtat_b200_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def _declare(L):
    vp, i, sz, u64 = C.c_void_p, C.c_int, C.c_size_t, C.c_uint64
    dp, fp = C.POINTER(C.c_double), C.POINTER(C.c_float)
    L.rtat_b200_version.restype = i
    L.rtat_b200_status_string.restype = C.c_char_p
    L.rtat_b200_status_string.argtypes = [i]
    L.rtat_b200_host_schedule.argtypes = [i, i, i, i, i, i, i, i, i, C.POINTER(i), i]
    L.rtat_b200_create.argtypes = [C.POINTER(vp)]
    L.rtat_b200_destroy.argtypes = [vp]
    L.rtat_b200_set_stream.argtypes = [vp, vp]
    L.rtat_b200_get_stream.argtypes = [vp, C.POINTER(vp)]
    for name, sp in (("d", dp), ("s", fp)):
        getattr(L, f"rtat_b200_{name}gemm").argtypes = [vp, i, i, i, i, i, sp, vp, i, vp, i, sp, vp, i]
        getattr(L, f"rtat_b200_{name}gemm_host").argtypes = [vp, i, i, i, i, i, sp, vp, i, vp, i, sp, vp, i]
        getattr(L, f"rtat_b200_{name}geam").argtypes = [vp, i, i, i, i, sp, vp, i, sp, vp, i, vp, i]
        getattr(L, f"rtat_b200_{name}syrk").argtypes = [vp, i, i, i, i, sp, vp, i, sp, vp, i]
        getattr(L, f"rtat_b200_{name}trsm").argtypes = [vp, i, i, i, i, i, i, sp, vp, i, vp, i]
    L.rtat_b200_fill_uniform_d.argtypes = [vp, vp, sz, u64, C.c_double, C.c_double]
    L.rtat_b200_fill_uniform_s.argtypes = [vp, vp, sz, u64, C.c_float, C.c_float]
    L.rtat_b200_set_option.argtypes = [vp, C.c_char_p, i]
    L.rtat_b200_get_option.argtypes = [vp, C.c_char_p, C.POINTER(i)]
    L.rtat_b200_last_kernel.restype = C.c_char_p
    L.rtat_b200_last_kernel.argtypes = [vp]
    L.rtat_b200_launch_count.restype = u64
    L.rtat_b200_launch_count.argtypes = [vp]


def _check(status, what):
    if status != 0:
        name = lib().rtat_b200_status_string(status).decode()
        raise RtatError(f"{what} failed: {name} ({status})")


def _ptr(x):
    """Device (or host) address of a torch tensor / numpy array / int."""
    if x is None:
        return None
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    return x.ctypes.data


class Handle:
    """RAII wrapper of rtat_b200_handle_t.  Mirrors how the reference's drivers hold a
    cublasHandle_t bound to one stream (src/app/test_harness.h:147-155)."""

    def __init__(self, stream=None):
        self._h = C.c_void_p()
        _check(lib().rtat_b200_create(C.byref(self._h)), "rtat_b200_create")
        if stream is not None:
            self.set_stream(stream)

    def close(self):
        if self._h:
            lib().rtat_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def raw(self):
        return self._h

    def set_stream(self, stream):
        s = getattr(stream, "cuda_stream", stream)
        _check(lib().rtat_b200_set_stream(self._h, C.c_void_p(s)), "rtat_b200_set_stream")

    def get_stream(self):
        out = C.c_void_p()
        _check(lib().rtat_b200_get_stream(self._h, C.byref(out)), "rtat_b200_get_stream")
        return out.value or 0

    def set_option(self, key, value):
        _check(lib().rtat_b200_set_option(self._h, key.encode(), int(value)), "rtat_b200_set_option")

    def last_kernel(self):
        return lib().rtat_b200_last_kernel(self._h).decode()

    def launch_count(self):
        return int(lib().rtat_b200_launch_count(self._h))

    # ---- BLAS entry points: thin, status-checked; dtype picks d/s --------------------------------
    @staticmethod
    def _scal(double, v):
        return C.byref(C.c_double(v)) if double else C.byref(C.c_float(v))

    def gemm(self, double, transa, transb, m, n, k, alpha, A, lda, B, ldb, beta, Cm, ldc, host=False, check=True):
        fn = getattr(lib(), f"rtat_b200_{'d' if double else 's'}gemm{'_host' if host else ''}")
        st = fn(self._h, transa, transb, m, n, k, self._scal(double, alpha), _ptr(A), lda, _ptr(B), ldb,
                self._scal(double, beta), _ptr(Cm), ldc)
        if check:
            _check(st, fn.__name__)
        return st

    def geam(self, double, transa, transb, m, n, alpha, A, lda, beta, B, ldb, Cm, ldc, check=True):
        fn = getattr(lib(), f"rtat_b200_{'d' if double else 's'}geam")
        st = fn(self._h, transa, transb, m, n, self._scal(double, alpha), _ptr(A), lda,
                self._scal(double, beta), _ptr(B), ldb, _ptr(Cm), ldc)
        if check:
            _check(st, fn.__name__)
        return st

    def syrk(self, double, uplo, trans, n, k, alpha, A, lda, beta, Cm, ldc, check=True):
        """uplo: 0 lower / 1 upper (RTAT_B200_FILL_*); trans: 0 N / 1 T."""
        fn = getattr(lib(), f"rtat_b200_{'d' if double else 's'}syrk")
        st = fn(self._h, uplo, trans, n, k, self._scal(double, alpha), _ptr(A), lda, self._scal(double, beta), _ptr(Cm), ldc)
        if check:
            _check(st, fn.__name__)
        return st

    def trsm(self, double, side, uplo, trans, diag, m, n, alpha, A, lda, B, ldb, check=True):
        """side: 0 left / 1 right; uplo: 0 lower / 1 upper; trans: 0 N / 1 T; diag: 0 non-unit / 1 unit."""
        fn = getattr(lib(), f"rtat_b200_{'d' if double else 's'}trsm")
        st = fn(self._h, side, uplo, trans, diag, m, n, self._scal(double, alpha), _ptr(A), lda, _ptr(B), ldb)
        if check:
            _check(st, fn.__name__)
        return st

    def fill_uniform_ptr(self, ptr, n, seed, lo=-1.0, hi=1.0, double=True):
        fn = lib().rtat_b200_fill_uniform_d if double else lib().rtat_b200_fill_uniform_s
        _check(fn(self._h, ptr, n, seed, lo, hi), "fill_uniform")

    def fill_uniform(self, x, seed, lo=-1.0, hi=1.0):
        import torch

        if x.dtype == torch.float64:
            _check(lib().rtat_b200_fill_uniform_d(self._h, x.data_ptr(), x.numel(), seed, lo, hi), "fill_uniform_d")
        elif x.dtype == torch.float32:
            _check(lib().rtat_b200_fill_uniform_s(self._h, x.data_ptr(), x.numel(), seed, lo, hi), "fill_uniform_s")
        else:
            raise RtatError("fill_uniform: float32/float64 only")
