"""Transposing geam: the register-blocked kernel (geam.transpose=2) against the TMA pipeline (1) and the scalar tile kernel (3):
bit-exactness against torch on aligned / odd / offset operands with and without beta, then GB/s.  usage: geam_transpose.py [check|perf|all]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt

h = rt.Handle(stream=torch.cuda.current_stream())
what = sys.argv[1] if len(sys.argv) > 1 else "all"


def fma_ref_ok(dtype, m, n, lda_pad, ldc_pad, a_off, c_off, alpha, beta, sel, order=0):
    """beta != 0: compare against the scalar tile kernel (geam.transpose=3), which the goldens pin to cuBLAS"""
    dbl = dtype == torch.float64
    lda, ldc = n + lda_pad, m + ldc_pad
    g = torch.Generator(device="cuda").manual_seed(m * 131 + n)
    bufA = torch.rand(lda * m + 4, dtype=dtype, device="cuda", generator=g) * 2 - 1
    bufC = torch.rand(ldc * n + 4, dtype=dtype, device="cuda", generator=g) * 2 - 1
    outs = []
    for s_ in (sel, 3):
        A, Cm = bufA[a_off:], bufC.clone()[c_off:]
        h.set_option("geam.transpose", s_)
        h.set_option("geam.order", order)
        h.geam(dbl, 1, 0, m, n, alpha, A, lda, beta, Cm, ldc, Cm, ldc)
        torch.cuda.synchronize()
        outs.append((Cm, h.last_kernel()))
    return bool(torch.equal(outs[0][0], outs[1][0])), outs[0][1] + " vs " + outs[1][1]


ok = True
if what in ("check", "all"):
    for dtype in (torch.float64, torch.float32):
        e = 2 if dtype == torch.float64 else 4
        for (m, n) in [(64, 64), (128, 256), (517, 300), (1000, 999), (63, 5), (2049, 4097), (4096, 130)]:
            for lda_pad, ldc_pad, a_off, c_off in [(0, 0, 0, 0), (1, 0, 0, 0), (0, 3, 0, 0), (e, e, 1, 0), (0, 0, 0, 1), (1, 1, 1, 1)]:
                for order in (0, 1):
                    good, name = fma_ref_ok(dtype, m, n, lda_pad, ldc_pad, a_off, c_off, 1.0, 0.0, 2, order)
                    good2, name2 = fma_ref_ok(dtype, m, n, lda_pad, ldc_pad, a_off, c_off, -1.5, 0.75, 2, order)
                    if not (good and good2):
                        print("MISMATCH", dtype, m, n, lda_pad, ldc_pad, a_off, c_off, order, good, good2, name, name2, flush=True)
                    ok &= good and good2
        print(dtype, "checked; last kernels:", name, "|", name2, flush=True)
    print("CHECK", "PASS" if ok else "FAIL", flush=True)


def timeit(fn, reps=7):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            fn()
        b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b) / 10)
    return best


if what in ("perf", "all"):
    shapes = [("8192^2 f64", torch.float64, 8192, 8192, 0, 0), ("4097x2049 -> ld 2080 f64", torch.float64, 2049, 4097, 0, 31), ("2049x4097 -> tight odd ld f64", torch.float64, 4097, 2049, 0, 0),
              ("3001x2049 -> ld 2080", torch.float64, 2049, 3001, 0, 31), ("131072x128 f32", torch.float32, 128, 131072, 0, 0), ("8192^2 f32", torch.float32, 8192, 8192, 0, 0),
              ("16384^2 f64", torch.float64, 16384, 16384, 0, 0)]
    for name, dtype, m, n, lda_pad, ldc_pad in shapes:
        lda, ldc = n + lda_pad, m + ldc_pad
        A = torch.empty(lda * m, dtype=dtype, device="cuda"); h.fill_uniform(A, 1)
        Cm = torch.empty(ldc * n, dtype=dtype, device="cuda")
        for sel, order in ((1, 0), (2, 0), (2, 1), (3, 0)):
            h.set_option("geam.transpose", sel); h.set_option("geam.order", order)
            try:
                ms = timeit(lambda: h.geam(dtype == torch.float64, 1, 0, m, n, 1.0, A, lda, 0.0, Cm, ldc, Cm, ldc))
            except rt.RtatError as ex:
                print(name, sel, "unsupported", ex); continue
            print(f"{name:32s} sel={sel} order={order} {h.last_kernel():36s} {ms*1e3:8.1f} us {2.0*m*n*A.element_size()/ms*1e-6:8.1f} GB/s", flush=True)
    h.set_option("geam.transpose", 0); h.set_option("geam.order", 0)
    for name, dtype, m, n, lda_pad, ldc_pad in shapes[:2] + shapes[4:6]:
        lda, ldc = n + lda_pad, m + ldc_pad
        A = torch.empty(lda * m, dtype=dtype, device="cuda"); h.fill_uniform(A, 1)
        Cm = torch.empty(ldc * n, dtype=dtype, device="cuda")
        for upc in (0, 2):
            h.set_option("geam.transpose", 2); h.set_option("geam.upc", upc)
            ms = timeit(lambda: h.geam(dtype == torch.float64, 1, 0, m, n, 1.0, A, lda, 0.0, Cm, ldc, Cm, ldc))
            print(f"{name:32s} upc={upc} {h.last_kernel():36s} {ms*1e3:8.1f} us {2.0*m*n*A.element_size()/ms*1e-6:8.1f} GB/s", flush=True)
    h.set_option("geam.transpose", 0); h.set_option("geam.upc", 0)
    for name, m, n, ldc in [("pad-copy 4097x2049 -> ld 4128", 4097, 2049, 4128), ("pad-copy 3001x2049 -> ld 3008", 3001, 2049, 3008), ("copy 4096x2048 (vec16 path: ld 4098)", 4096, 2048, 4098)]:
        A = torch.empty(m * n, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
        Cm = torch.empty(ldc * n, dtype=torch.float64, device="cuda")
        for pers in (0, 1):
            h.set_option("geam.persistent", pers)
            ms = timeit(lambda: h.geam(True, 0, 0, m, n, 1.0, A, m, 0.0, Cm, ldc, Cm, ldc))
            print(f"{name:40s} persistent={pers} {h.last_kernel():24s} {ms*1e3:8.1f} us {2.0*m*n*8/ms*1e-6:8.1f} GB/s", flush=True)
    h.set_option("geam.persistent", 0)
    # copy tuning
    for name, dtype, m, n in [("copy 8192^2 f64", torch.float64, 8192, 8192), ("copy 8192^2 f32", torch.float32, 8192, 8192), ("copy 16384^2 f64", torch.float64, 16384, 16384)]:
        A = torch.empty(m * n, dtype=dtype, device="cuda"); h.fill_uniform(A, 1)
        Cm = torch.empty(m * n, dtype=dtype, device="cuda")
        for t in (0, 12, 8):
            h.set_option("geam.tune", t)
            ms = timeit(lambda: h.geam(dtype == torch.float64, 0, 0, m, n, 1.0, A, m, 0.0, Cm, m, Cm, m))
            print(f"{name:24s} tune={t:2d} {h.last_kernel():24s} {ms*1e3:8.1f} us {2.0*m*n*A.element_size()/ms*1e-6:8.1f} GB/s", flush=True)
        h.set_option("geam.tune", 0)
sys.exit(0 if ok else 1)
