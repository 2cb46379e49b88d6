"""GEMV-shaped DGEMM calls: skinny kernel (dgemm.skinny = 2) against the tile kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
from perf_tiles3 import timeit, h


def case(name, ta, tb, m, n, k, lda=None, ldb=None):
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    lda, ldb = lda or ar, ldb or br
    A = torch.empty(lda * ac, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(ldb * bc, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.zeros(m * n, dtype=torch.float64, device="cuda")
    ref = None
    big = 8.0 * (ar * ac + br * bc)
    for sk in (1, 2):
        h.set_option("dgemm.skinny", sk)
        ms = timeit(lambda: h.gemm(True, ta, tb, m, n, k, 1.0, A, lda, B, ldb, 0.0, Cm, m))
        torch.cuda.synchronize()
        if ref is None:
            ref, diff = Cm.clone(), 0.0
        else:
            diff = float(((Cm - ref).abs().max() / ref.abs().max()).item())
        print(f"{name:40s} skinny {sk} {h.last_kernel():40s} {ms*1e3:9.2f} us {big/ms*1e-6:8.1f} GB/s  maxdiff {diff:.1e}", flush=True)
    h.set_option("dgemm.skinny", 1)


for r in (1, 2, 4, 8, 16):
    case(f"m-strip r={r} TT 3001x2049 (ld 2080, 3008)", 1, 1, r, 3001, 2049, lda=2080, ldb=3008)
for r in (1, 4, 8, 16):
    case(f"m-strip r={r} NN n=8192 k=2048 (X K-contig)", 0, 0, r, 8192, 2048, lda=8200)
for r in (1, 4, 8, 16):
    case(f"n-strip r={r} NN m=8192 k=2048 (X M-contig)", 0, 0, 8192, r, 2048)
for r in (1, 8):
    case(f"n-strip r={r} TN m=8192 k=2048 (X K-contig)", 1, 0, 8192, r, 2048)
    case(f"m-strip r={r} NT n=8192 k=4096 (X N-contig)", 0, 1, r, 8192, 4096, lda=8200)
