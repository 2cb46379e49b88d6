"""Peeled strips: skinny kernel against the 64x64-tile strip launch, on config 3's shapes and a few others."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
from perf_tiles3 import timeit, h  # noqa: E402  (runs that probe's cases on import: keep this file's own list short)


def case(name, ta, tb, m, n, k, lda=None, ldb=None, tile=0):
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    lda, ldb = lda or ar, ldb or br
    A = torch.empty(lda * ac, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(ldb * bc, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.zeros(m * n, dtype=torch.float64, device="cuda")
    ref = None
    for sk in (0, 1):
        h.set_option("dgemm.tile", tile)
        h.set_option("dgemm.skinny", sk)
        ms = timeit(lambda: h.gemm(True, ta, tb, m, n, k, 1.0, A, lda, B, ldb, 0.0, Cm, m))
        torch.cuda.synchronize()
        if ref is None:
            ref, diff = Cm.clone(), 0.0
        else:
            diff = float(((Cm - ref).abs().max() / ref.abs().max()).item())
        print(f"{name:44s} skinny {sk} {h.last_kernel():56s} {ms:9.4f} ms {2.0*m*n*k/ms*1e-9:7.2f} TF  maxdiff {diff:.1e}", flush=True)
    h.set_option("dgemm.tile", 0)
    h.set_option("dgemm.skinny", 1)


case("C3 pad plan TT 4097x3001x2049 (2080, 3008)", 1, 1, 4097, 3001, 2049, lda=2080, ldb=3008)
case("C3 pad plan NT (4128, 3008)", 0, 1, 4097, 3001, 2049, lda=4128, ldb=3008)
case("C3 NT unpadded (cp.async)", 0, 1, 4097, 3001, 2049)
case("NN 8193x8192x4096", 0, 0, 8193, 8192, 4096)
case("TN 8193x8192x4096", 1, 0, 8193, 8192, 4096)
case("NN 4100x4104x4096", 0, 0, 4100, 4104, 4096)
case("TT 4100x4104x4096", 1, 1, 4100, 4104, 4096)
case("NN 8200x8200x2048", 0, 0, 8200, 8200, 2048)
case("TN 8208x8208x2048 (16-wide strips)", 1, 0, 8208, 8208, 2048)
