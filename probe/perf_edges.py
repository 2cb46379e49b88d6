"""Ragged-tile experiment for the persistent DGEMM: config-3-like shapes, tile 128 / 64, both the TMA path
(padded leading dimensions) and the cp.async path (tight odd leading dimensions); each result checked against torch.matmul."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt

h = rt.Handle(stream=torch.cuda.current_stream())


def timeit(fn, reps=7, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(4):
            fn()
        b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b) / 4)
    return best


def case(ta, tb, m, n, k, pad):
    up = (lambda x: -(-x // 32) * 32) if pad else (lambda x: x)
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    lda, ldb, ldc = up(ar), up(br), m
    A = torch.zeros(lda * ac, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    B = torch.zeros(ldb * bc, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.empty(ldc * n, dtype=torch.float64, device="cuda")
    Av = A.view(ac, lda)[:, :ar]; Bv = B.view(bc, ldb)[:, :br]          # row-major views = stored^T
    opAT = Av if not ta else Av.t()                                     # (k x m)
    opBT = Bv if not tb else Bv.t()                                     # (n x k)
    ref = opBT @ opAT                                                   # C^T (n x m)
    bound = opBT.abs() @ opAT.abs()
    for tile in (128, 64):
        for w in (0,):
            h.set_option("dgemm.tile", tile)
            Cm.fill_(float("nan"))
            h.gemm(True, ta, tb, m, n, k, 1.0, A, lda, B, ldb, 0.0, Cm, ldc)
            torch.cuda.synchronize()
            err = (Cm.view(n, ldc)[:, :m] - ref).abs()
            ok = bool((err <= 2 * k * 2.2e-16 * bound + 1e-300).all())
            ms = timeit(lambda: h.gemm(True, ta, tb, m, n, k, 1.0, A, lda, B, ldb, 0.0, Cm, ldc))
            print(f"{'TN'[1-ta]}{'TN'[1-tb]} {m}x{n}x{k} pad={pad} tile {tile}: {ms:8.4f} ms {2.0*m*n*k/ms*1e-9:7.2f} TF  {h.last_kernel()}  {'ok' if ok else 'MISMATCH max err %.3e' % err.max().item()}", flush=True)
    h.set_option("dgemm.tile", 0)
    ms = timeit(lambda: h.gemm(True, ta, tb, m, n, k, 1.0, A, lda, B, ldb, 0.0, Cm, ldc))
    print(f"   auto: {ms:8.4f} ms {2.0*m*n*k/ms*1e-9:7.2f} TF  {h.last_kernel()}", flush=True)


case(0, 1, 4097, 3001, 2049, True)     # config 3 after the pad plan (TMA producers)
case(0, 0, 4097, 3001, 2049, True)
case(0, 1, 4097, 3001, 2049, False)    # config 3 in place (cp.async producers)
case(1, 0, 4097, 3001, 2049, True)
case(0, 0, 4100, 4100, 4096, True)     # 4 rows / columns past a tile boundary
case(0, 0, 8192, 8192, 2048, True)     # no ragged tiles: schedule must be unchanged
case(1, 1, 1000, 777, 555, True)
case(0, 0, 130, 129, 4096, True)
case(1, 0, 8192, 8192, 8192, True)     # config 2
case(0, 0, 8192, 8192, 1024, True)     # K-panel of the host pipeline
case(0, 0, 8192, 1024, 8192, True)     # column block of the host pipeline
case(0, 0, 1024, 1024, 1024, True)     # config 1
case(0, 0, 2048, 2048, 2048, True)
case(0, 0, 4096, 4096, 4096, True)
