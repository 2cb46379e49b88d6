"""ctypes binding of include/rtat_b200_mg.h: column-split GEMM across the GPUs of one node."""
import ctypes as C

from ._capi import Handle, RtatError, _check, _ptr, lib

HANDLE_BYTES = 64
_declared = False


def _mg_lib():
    global _declared
    L = lib()
    if not _declared:
        vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
        ip = C.POINTER(C.c_int)
        L.rtat_b200_mg_partition.argtypes = [i, i, i, ip, ip]
        L.rtat_b200_mg_panel.argtypes = [i, i, i, ip, ip]
        L.rtat_b200_mg_auto_panel.argtypes = [i, i, ip, ip]
        L.rtat_b200_mg_create.argtypes = [C.POINTER(vp), vp, i, i]
        L.rtat_b200_mg_destroy.argtypes = [vp]
        L.rtat_b200_mg_alloc_shared.argtypes = [vp, sz, C.POINTER(vp), C.c_char_p]
        L.rtat_b200_mg_free_shared.argtypes = [vp, vp]
        L.rtat_b200_mg_open_shared.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
        L.rtat_b200_mg_close_shared.argtypes = [vp, vp]
        dp = C.POINTER(C.c_double)
        L.rtat_b200_mg_dgemm.argtypes = [vp, i, i, i, i, i, dp, vp, i, vp, vp, i, dp, vp, i, i]
        L.rtat_b200_mg_host_setup.argtypes = [vp, sz, C.c_char_p]
        L.rtat_b200_mg_host_connect.argtypes = [vp, C.c_char_p]
        L.rtat_b200_mg_host_teardown.argtypes = [vp]
        L.rtat_b200_mg_dgemm_host.argtypes = [vp, i, i, i, i, i, dp, vp, i, vp, i, dp, vp, i]
        _declared = True
    return L


def partition(n, world, rank):
    """Rank's contiguous column block (first, count) of n columns.  Pure host logic (no GPU needed)."""
    f, c = C.c_int(), C.c_int()
    _check(_mg_lib().rtat_b200_mg_partition(n, world, rank, C.byref(f), C.byref(c)), "rtat_b200_mg_partition")
    return f.value, c.value


def panel(k, panels, p):
    """K-panel p of `panels` over extent k: (first, count).  Pure host logic."""
    f, c = C.c_int(), C.c_int()
    _check(_mg_lib().rtat_b200_mg_panel(k, panels, p, C.byref(f), C.byref(c)), "rtat_b200_mg_panel")
    return f.value, c.value


def auto_panel(k, p):
    """Panel p (0..3) of the automatic schedule: (first, count).  Pure host logic."""
    f, c = C.c_int(), C.c_int()
    _check(_mg_lib().rtat_b200_mg_auto_panel(k, p, C.byref(f), C.byref(c)), "rtat_b200_mg_auto_panel")
    return f.value, c.value


class ColumnSplit:
    def __init__(self, handle: Handle, rank, world):
        self.rank, self.world, self.handle = rank, world, handle
        self._mg = C.c_void_p()
        _check(_mg_lib().rtat_b200_mg_create(C.byref(self._mg), handle.raw, rank, world), "rtat_b200_mg_create")

    def close(self):
        if self._mg:
            _mg_lib().rtat_b200_mg_destroy(self._mg)
            self._mg = C.c_void_p()

    def alloc_shared(self, nbytes):
        ptr = C.c_void_p()
        buf = C.create_string_buffer(HANDLE_BYTES)
        _check(_mg_lib().rtat_b200_mg_alloc_shared(self._mg, nbytes, C.byref(ptr), buf), "rtat_b200_mg_alloc_shared")
        return ptr.value, buf.raw

    def free_shared(self, ptr):
        _check(_mg_lib().rtat_b200_mg_free_shared(self._mg, C.c_void_p(ptr)), "rtat_b200_mg_free_shared")

    def open_shared(self, handle_bytes):
        ptr = C.c_void_p()
        st = _mg_lib().rtat_b200_mg_open_shared(self._mg, bytes(handle_bytes), C.byref(ptr))
        if st != 0:
            raise RtatError(f"rtat_b200_mg_open_shared failed ({st}): peer mapping unavailable")
        return ptr.value

    def close_shared(self, ptr):
        _mg_lib().rtat_b200_mg_close_shared(self._mg, C.c_void_p(ptr))

    def dgemm(self, transa, transb, m, n_loc, k, alpha, A_src, lda, A_stage, B_loc, ldb, beta, C_loc, ldc, panels=0):
        al, be = C.c_double(alpha), C.c_double(beta)
        _check(_mg_lib().rtat_b200_mg_dgemm(self._mg, transa, transb, m, n_loc, k, C.byref(al), _ptr(A_src), lda, _ptr(A_stage),
                                            _ptr(B_loc), ldb, C.byref(be), _ptr(C_loc), ldc, panels), "rtat_b200_mg_dgemm")

    # ---- host-buffer column split (A, B_loc, C_loc in pinned host memory) ----
    def host_setup(self, bytes_a):
        """Allocate this rank's replica of A (+ flag page) and return its 64-byte IPC handle for the all-gather."""
        buf = C.create_string_buffer(HANDLE_BYTES)
        _check(_mg_lib().rtat_b200_mg_host_setup(self._mg, bytes_a, buf), "rtat_b200_mg_host_setup")
        return buf.raw

    def host_connect(self, handles):
        """handles: world x 64 bytes in rank order (None when world == 1)."""
        hb = None if handles is None else bytes(handles)
        if hb is not None and len(hb) != HANDLE_BYTES * self.world:
            raise RtatError("host_connect needs world x 64 handle bytes")
        st = _mg_lib().rtat_b200_mg_host_connect(self._mg, hb)
        if st != 0:
            raise RtatError(f"rtat_b200_mg_host_connect failed ({st}): peer mapping unavailable")

    def host_teardown(self):
        _mg_lib().rtat_b200_mg_host_teardown(self._mg)

    def dgemm_host(self, transa, transb, m, n_loc, k, alpha, A, lda, B_loc, ldb, beta, C_loc, ldc):
        """Collective: every rank passes the full host A and its own column block of op(B) / C; returns when C_loc landed."""
        al, be = C.c_double(alpha), C.c_double(beta)
        _check(_mg_lib().rtat_b200_mg_dgemm_host(self._mg, transa, transb, m, n_loc, k, C.byref(al), _ptr(A), lda, _ptr(B_loc), ldb,
                                                 C.byref(be), _ptr(C_loc), ldc), "rtat_b200_mg_dgemm_host")
