"""CTA-pair (cta_group::2) SGEMM against the single-CTA kernel and an FP64 product: bit-equality between the two kernels
(same products, same accumulation order), the north-star bound against FP64, and timings.  usage: sgemm_pair.py [check|perf|all]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt

h = rt.Handle(stream=torch.cuda.current_stream())
what = sys.argv[1] if len(sys.argv) > 1 else "all"
EPS = 2.0 ** -23


def run(ta, tb, m, n, k, pair, alpha=1.0, beta=0.0, lda_pad=0, ldb_pad=0, ldc_pad=0, seed=3):
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    lda, ldb, ldc = ar + lda_pad, br + ldb_pad, m + ldc_pad
    g = torch.Generator(device="cuda").manual_seed(seed)
    A = torch.rand(lda * ac, dtype=torch.float32, device="cuda", generator=g) * 2 - 1
    B = torch.rand(ldb * bc, dtype=torch.float32, device="cuda", generator=g) * 2 - 1
    C0 = torch.rand(ldc * n, dtype=torch.float32, device="cuda", generator=g) * 2 - 1
    C = C0.clone()
    h.set_option("sgemm.variant", 2)
    h.set_option("sgemm.pair", pair)
    h.gemm(False, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
    torch.cuda.synchronize()
    return A, B, C0, C, (lda, ldb, ldc), h.last_kernel()


def check(ta, tb, m, n, k, **kw):
    A, B, C0, C1, (lda, ldb, ldc), k1 = run(ta, tb, m, n, k, 1, **kw)
    _, _, _, C2, _, k2 = run(ta, tb, m, n, k, 0, **kw)
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    Am = A.view(ac, lda)[:, :ar].double()
    Bm = B.view(bc, ldb)[:, :br].double()
    opA = Am if ta else Am.t()          # (m, k)
    opB = Bm if tb else Bm.t()
    alpha, beta = kw.get("alpha", 1.0), kw.get("beta", 0.0)
    want = alpha * (opA @ opB) + beta * C0.view(n, ldc)[:, :m].t().double()
    bound = 8 * k * EPS * (opA.abs() @ opB.abs()) * abs(alpha) + EPS * want.abs()
    got1 = C1.view(n, ldc)[:, :m].t().double()
    got2 = C2.view(n, ldc)[:, :m].t().double()
    r1 = ((got1 - want).abs() / bound).max().item()
    r2 = ((got2 - want).abs() / bound).max().item()
    same = bool(torch.equal(C1, C2))
    pad_ok = bool(torch.equal(C1.view(n, ldc)[:, m:], C0.view(n, ldc)[:, m:]))
    print(f"op {'T' if ta else 'N'}{'T' if tb else 'N'} {m}x{n}x{k} {kw}: pair err/bound {r1:.3f} single {r2:.3f} bit-equal {same} padding untouched {pad_ok}  [{k1}]", flush=True)
    return r1 <= 1.0 and same and pad_ok


def timeit(fn, reps=5):
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


ok = True
if what in ("check", "all"):
    for ta in (0, 1):
        for tb in (0, 1):
            ok &= check(ta, tb, 512, 256, 256)
            ok &= check(ta, tb, 1000, 300, 520, alpha=1.5, beta=-0.5, ldc_pad=3)
            ok &= check(ta, tb, 384, 128, 100)        # odd number of M-tiles: the peer's half of the last pair is empty
    ok &= check(0, 0, 4096, 128, 4096)
    ok &= check(1, 0, 2048, 2048, 1024, beta=1.0)
    print("CHECK", "PASS" if ok else "FAIL", flush=True)
if what in ("perf", "all"):
    for name, ta, tb, m, n, k in [("C4 131072x128x4096", 0, 0, 131072, 128, 4096), ("4096^3 NN", 0, 0, 4096, 4096, 4096), ("4096^3 TN", 1, 0, 4096, 4096, 4096),
                                  ("8192^3 NT", 0, 1, 8192, 8192, 8192)]:
        ar, ac = (k, m) if ta else (m, k)
        br, bc = (n, k) if tb else (k, n)
        A = torch.empty(ar * ac, dtype=torch.float32, device="cuda"); h.fill_uniform(A, 1)
        B = torch.empty(br * bc, dtype=torch.float32, device="cuda"); h.fill_uniform(B, 2)
        C = torch.empty(m * n, dtype=torch.float32, device="cuda")
        for pair in (0, 1):
            h.set_option("sgemm.pair", pair)
            ms = timeit(lambda: h.gemm(False, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, C, m))
            print(f"{name:22s} pair={pair} {h.last_kernel():48s} {ms:8.4f} ms {2.0*m*n*k/ms*1e-9:8.1f} TFLOP/s", flush=True)
sys.exit(0 if ok else 1)
