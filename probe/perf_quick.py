"""Quick kernel timing through the C ABI (not the bench contract; development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt

h = rt.Handle(stream=torch.cuda.current_stream())


def timeit(fn, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); b.synchronize()
    batch = max(1, min(50, int(2.0 / max(a.elapsed_time(b), 1e-3))))  # keep the queue full for short kernels
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(batch):
            fn()
        b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b) / batch)
    return best


def dgemm_cfg(name, ta, tb, m, n, k, variants=(2, 3, 4), dtype=torch.float64, opts=()):
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    A = torch.empty(ar * ac, dtype=dtype, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(br * bc, dtype=dtype, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.empty(m * n, dtype=dtype, device="cuda")
    dbl = dtype == torch.float64
    key = "dgemm.variant" if dbl else "sgemm.variant"
    for v in variants:
        h.set_option(key, v)
        for k_, v_ in opts: h.set_option(k_, v_)
        try:
            ms = timeit(lambda: h.gemm(dbl, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, Cm, m))
            print(f"{name:34s} variant {v} {h.last_kernel():36s} {ms:9.3f} ms {2.0*m*n*k/ms*1e-9:8.2f} TFLOP/s", flush=True)
        except rt.RtatError as e:
            print(f"{name:34s} variant {v} unsupported: {e}")
        h.set_option(key, 0)


def geam_cfg(name, ta, m, n, ldc=None, dtype=torch.float64):
    ar, ac = (n, m) if ta else (m, n)
    ldc = ldc or m
    A = torch.empty(ar * ac, dtype=dtype, device="cuda"); h.fill_uniform(A, 1)
    Cm = torch.empty(ldc * n, dtype=dtype, device="cuda")
    ms = timeit(lambda: h.geam(dtype == torch.float64, ta, 0, m, n, 1.0, A, ar, 0.0, Cm, ldc, Cm, ldc))
    print(f"{name:34s} {h.last_kernel():30s} {ms:9.4f} ms {2.0*m*n*A.element_size()/ms*1e-6:8.1f} GB/s", flush=True)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which == "tiles":
    for t in (128, 64):
        h.set_option("dgemm.tile", t)
        print("tile", t)
        dgemm_cfg("C1 dgemm NN 1024^3", 0, 0, 1024, 1024, 1024, variants=(3,))
        dgemm_cfg("C2 dgemm TN 8192^3", 1, 0, 8192, 8192, 8192, variants=(3,))
        dgemm_cfg("C3 dgemm NT 4097x3001x2049", 0, 1, 4097, 3001, 2049, variants=(3,))
        dgemm_cfg("dgemm NN 2048^3", 0, 0, 2048, 2048, 2048, variants=(3,))
        dgemm_cfg("dgemm NN 4096^3", 0, 0, 4096, 4096, 4096, variants=(3,))
        dgemm_cfg("C5/8-like NN 8192x1024x8192", 0, 0, 8192, 1024, 8192, variants=(3,))
        dgemm_cfg("dgemm NN 512^3", 0, 0, 512, 512, 512, variants=(3,))
    h.set_option("dgemm.tile", 0)
if which in ("all", "dgemm"):
    dgemm_cfg("C1 dgemm NN 1024^3", 0, 0, 1024, 1024, 1024)
    dgemm_cfg("C2 dgemm TN 8192^3", 1, 0, 8192, 8192, 8192, variants=(1, 3, 4))
    dgemm_cfg("C2' dgemm NN 8192^3", 0, 0, 8192, 8192, 8192, variants=(1, 3, 4))
    dgemm_cfg("C2'' dgemm NT 8192^3", 0, 1, 8192, 8192, 8192, variants=(1, 3, 4))
    dgemm_cfg("C2''' dgemm TT 8192^3", 1, 1, 8192, 8192, 8192, variants=(1, 3, 4))
    dgemm_cfg("C3 dgemm NT 4097x3001x2049", 0, 1, 4097, 3001, 2049)
    dgemm_cfg("dgemm NN 4096^3", 0, 0, 4096, 4096, 4096)
    dgemm_cfg("dgemm NN 2048^3", 0, 0, 2048, 2048, 2048)
    dgemm_cfg("C5/8-like NN 8192x1024x8192", 0, 0, 8192, 1024, 8192)
if which == "c3":
    for split in (1, 0):
        h.set_option("dgemm.splitk_tma", split)
        for t in (0, 64, 128):
            h.set_option("dgemm.tile", t)
            print("splitk_tma", split, "tile", t, end="  ")
            dgemm_cfg("C3 dgemm NT 4097x3001x2049", 0, 1, 4097, 3001, 2049, variants=(3,))
    h.set_option("dgemm.tile", 0); h.set_option("dgemm.splitk_tma", 1)
    dgemm_cfg("dgemm NN 4097x4096x4096 (A odd ld)", 0, 0, 4097, 4096, 4096, variants=(3, 4))
    dgemm_cfg("dgemm NT 8191^3", 0, 1, 8191, 8191, 8191, variants=(3, 4))
if which == "beta":
    # K-panel products of the host pipeline: 8192 x 4096 x 1024 with beta = 1 (config 2's phase 1) and the TRSM update shape
    for name, ta, tb, m, n, k in [("panel TN 8192x4096x1024", 1, 0, 8192, 4096, 1024), ("panel NN 8192x4096x1024", 0, 0, 8192, 4096, 1024),
                                  ("update NN 1024x8192x1024", 0, 0, 1024, 8192, 1024), ("mg panel NN 32768x4096x2048", 0, 0, 32768, 4096, 2048)]:
        ar, ac = (k, m) if ta else (m, k)
        br, bc = (n, k) if tb else (k, n)
        A = torch.empty(ar * ac, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
        B = torch.empty(br * bc, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
        Cm = torch.zeros(m * n, dtype=torch.float64, device="cuda")
        for beta in (0.0, 1.0):
            for pf in (1, 0):
                h.set_option("dgemm.prefetch_c", pf)
                ms = timeit(lambda: h.gemm(True, ta, tb, m, n, k, 1.0, A, ar, B, br, beta, Cm, m))
                print(f"{name:30s} beta={beta} prefetch_c={pf} {h.last_kernel():36s} {ms:9.3f} ms {2.0*m*n*k/ms*1e-9:8.2f} TFLOP/s", flush=True)
    h.set_option("dgemm.prefetch_c", 1)
if which == "c3pad":
    # config 3 after the explicit pad (what plan ..NPPN hands the multiply): A 4097 x 2049 ld 4128, B 3001 x 2049 ld 3008
    m, n, k = 4097, 3001, 2049
    A = torch.empty(4128 * k, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(3008 * k, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.empty(m * n, dtype=torch.float64, device="cuda")
    for tile in (0, 64, 128):
        for peel in (1, 0):
            h.set_option("dgemm.tile", tile); h.set_option("dgemm.peel", peel)
            ms = timeit(lambda: h.gemm(True, 0, 1, m, n, k, 1.0, A, 4128, B, 3008, 0.0, Cm, m))
            print(f"C3 padded NT tile={tile} peel={peel} {h.last_kernel():52s} {ms:9.4f} ms {2.0*m*n*k/ms*1e-9:8.2f} TFLOP/s", flush=True)
    h.set_option("dgemm.tile", 0); h.set_option("dgemm.peel", 1)
    for name, ta, tb, mm, nn, kk in [("NN 8193x8192x4096", 0, 0, 8193, 8192, 4096), ("TN 4100x4104x4096", 1, 0, 4100, 4104, 4096), ("NN 8200^2 x 2048", 0, 0, 8200, 8200, 2048)]:
        ar, ac = (kk, mm) if ta else (mm, kk)
        ar2 = ar + (ar % 2)
        A = torch.empty(ar2 * ac, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
        B = torch.empty(kk * nn, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
        Cm = torch.empty(mm * nn, dtype=torch.float64, device="cuda")
        for peel in (1, 0):
            h.set_option("dgemm.peel", peel)
            ms = timeit(lambda: h.gemm(True, ta, tb, mm, nn, kk, 1.0, A, ar2, B, kk, 0.0, Cm, mm))
            print(f"{name:24s} peel={peel} {h.last_kernel():52s} {ms:9.4f} ms {2.0*mm*nn*kk/ms*1e-9:8.2f} TFLOP/s", flush=True)
    h.set_option("dgemm.peel", 1)
if which == "geam_tune":
    for t in range(7):
        h.set_option("geam.tune", t)
        print("tune", t, end=" ")
        geam_cfg("copy 8192^2 f64", 0, 8192, 8192)
    h.set_option("geam.tune", 0)
if which in ("all", "geam"):
    geam_cfg("transpose 8192^2 f64", 1, 8192, 8192)
    geam_cfg("copy 8192^2 f64", 0, 8192, 8192)
    geam_cfg("pad-copy 4097x2049 ->ld4128", 0, 4097, 2049, 4128)
    geam_cfg("transpose 4097x2049 ->ld2080", 1, 2049, 4097, 2080)
    geam_cfg("pad-copy 3001x2049 ->ld3008", 0, 3001, 2049, 3008)
    geam_cfg("transpose 131072x128 f32", 1, 128, 131072, dtype=torch.float32)
if which == "sgemm_debug":
    for dbg in (0, 1, 4, 5):
        h.set_option("sgemm.debug", dbg)
        print("debug", dbg)
        dgemm_cfg("C4 sgemm NN 131072x128x4096", 0, 0, 131072, 128, 4096, variants=(2,), dtype=torch.float32)
        dgemm_cfg("sgemm TN 4096^3", 1, 0, 4096, 4096, 4096, variants=(2,), dtype=torch.float32)
    h.set_option("sgemm.debug", 0)
if which in ("all", "sgemm"):
    dgemm_cfg("C4 sgemm NN 131072x128x4096", 0, 0, 131072, 128, 4096, variants=(1, 2), dtype=torch.float32)
    dgemm_cfg("sgemm NN 4096^3", 0, 0, 4096, 4096, 4096, variants=(1, 2), dtype=torch.float32)
