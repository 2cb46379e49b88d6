"""Quick SYRK / TRSM timings: our C ABI vs cuBLAS (torch's handle through ctypes on libcublas) on the same buffers."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

import rtatblas_b200 as rt

cublas = C.CDLL("libcublas.so.12")
cb = C.c_void_p()
assert cublas.cublasCreate_v2(C.byref(cb)) == 0
cublas.cublasSetStream_v2(cb, C.c_void_p(torch.cuda.current_stream().cuda_stream))
h = rt.Handle(stream=torch.cuda.current_stream())


def timeit(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts)


def syrk(dt, n, k, lower, trans):
    dbl = dt == torch.float64
    A = torch.empty(n * k, dtype=dt, device="cuda")
    Cm = torch.zeros(n * n, dtype=dt, device="cuda")
    h.fill_uniform(A, 1)
    lda = k if trans else n
    ct = C.c_double if dbl else C.c_float
    one, zero = ct(1), ct(0)
    f = cublas.cublasDsyrk_v2 if dbl else cublas.cublasSsyrk_v2
    ours = timeit(lambda: h.syrk(dbl, 0 if lower else 1, trans, n, k, 1.0, A, lda, 0.0, Cm, n))
    name = h.last_kernel()
    ref = timeit(lambda: f(cb, 0 if lower else 1, trans, n, k, C.byref(one), C.c_void_p(A.data_ptr()), lda, C.byref(zero), C.c_void_p(Cm.data_ptr()), n))
    fl = n * (n + 1.0) * k
    print(f"syrk {'d' if dbl else 's'} n={n} k={k} lower={lower} trans={trans}: ours {ours:.3f} ms ({fl / ours / 1e9:.1f} TF)  cublas {ref:.3f} ms ({fl / ref / 1e9:.1f} TF)  ratio {ref / ours:.2f}  [{name}]", flush=True)


def trsm(dt, m, n, left, lower, trans):
    dbl = dt == torch.float64
    na = m if left else n
    A = torch.empty(na * na, dtype=dt, device="cuda")
    h.fill_uniform(A, 2)
    A.view(na, na).mul_(1.0 / 64).diagonal().add_(2.0)
    B0 = torch.empty(m * n, dtype=dt, device="cuda")
    h.fill_uniform(B0, 3)
    B = B0.clone()
    ct = C.c_double if dbl else C.c_float
    one = ct(1)
    f = cublas.cublasDtrsm_v2 if dbl else cublas.cublasStrsm_v2

    def ours_fn():
        B.copy_(B0)
        h.trsm(dbl, 0 if left else 1, 0 if lower else 1, trans, 0, m, n, 1.0, A, na, B, m)

    def ref_fn():
        B.copy_(B0)
        f(cb, 0 if left else 1, 0 if lower else 1, trans, 0, m, n, C.byref(one), C.c_void_p(A.data_ptr()), na, C.c_void_p(B.data_ptr()), m)

    copy = timeit(lambda: B.copy_(B0))
    ours = timeit(ours_fn) - copy
    ref = timeit(ref_fn) - copy
    fl = 1.0 * na * na * (n if left else m)
    print(f"trsm {'d' if dbl else 's'} m={m} n={n} left={left} lower={lower} trans={trans}: ours {ours:.3f} ms ({fl / ours / 1e9:.1f} TF)  cublas {ref:.3f} ms ({fl / ref / 1e9:.1f} TF)  ratio {ref / ours:.2f}", flush=True)


for n, k in [(1024, 1024), (4096, 4096), (8192, 8192), (8192, 512)]:
    syrk(torch.float64, n, k, 1, 0)
syrk(torch.float64, 8192, 8192, 0, 1)
syrk(torch.float64, 4097, 2049, 1, 0)
syrk(torch.float32, 8192, 8192, 1, 0)
syrk(torch.float32, 8192, 8192, 0, 1)
for m, n in [(1024, 1024), (4096, 4096), (8192, 8192)]:
    trsm(torch.float64, m, n, 1, 1, 0)
trsm(torch.float64, 8192, 8192, 0, 0, 1)
trsm(torch.float64, 8192, 1024, 1, 0, 0)
trsm(torch.float32, 8192, 8192, 1, 1, 0)
for base in (64, 256):
    h.set_option("trsm.base", base)
    print(f"-- trsm.base = {base}")
    trsm(torch.float64, 1024, 1024, 1, 1, 0)
    trsm(torch.float64, 8192, 8192, 1, 1, 0)
    trsm(torch.float64, 8192, 8192, 0, 0, 1)
h.set_option("trsm.base", 0)
