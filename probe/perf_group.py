"""Grouped raster height of the persistent DGEMM (option dgemm.group): time on 8192^3 TN / NN and 32768x4096x4096."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
from perf_tiles3 import timeit, h

only = int(sys.argv[1]) if len(sys.argv) > 1 else None
for (ta, tb, m, n, k) in [(1, 0, 8192, 8192, 8192), (0, 0, 8192, 8192, 8192), (0, 0, 16384, 4096, 8192)]:
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    A = torch.empty(ar * ac, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(br * bc, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.zeros(m * n, dtype=torch.float64, device="cuda")
    for g in ([only] if only else [2, 4, 6, 8, 12, 16, 32]):
        h.set_option("dgemm.group", g)
        if only:
            h.gemm(True, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, Cm, m)
            torch.cuda.synchronize()
            continue
        ms = timeit(lambda: h.gemm(True, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, Cm, m), reps=3)
        print(f"{'NT'[ta]}{'NT'[tb]} {m}x{n}x{k} group {g:2d}: {ms:8.3f} ms {2.0*m*n*k/ms*1e-9:6.2f} TF", flush=True)
    del A, B, Cm
