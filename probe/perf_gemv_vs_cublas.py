"""GEMV-shaped DGEMM calls: our skinny kernel against cuBLAS (through torch.matmul on the same column-major data)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
from perf_tiles3 import timeit, h

for (ta, tb, m, n, k) in [(1, 1, 1, 3001, 2049), (0, 1, 1, 8192, 4096), (0, 0, 1, 8192, 2048), (0, 0, 8192, 1, 2048), (1, 0, 8192, 1, 2048),
                          (0, 0, 4, 8192, 2048), (0, 1, 4, 8192, 4096), (0, 0, 8192, 4, 2048), (0, 0, 32768, 1, 32768), (1, 0, 32768, 2, 32768)]:
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    A = torch.empty(ar * ac, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(br * bc, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.zeros(m * n, dtype=torch.float64, device="cuda")
    ours = timeit(lambda: h.gemm(True, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, Cm, m))
    name = h.last_kernel()
    # column-major X (r x c, ld r) is torch's X.view(c, r).t(); C = op(A) op(B) column-major == (op(B)^T op(A)^T) row-major
    Am = A.view(ac, ar).t(); Bm = B.view(bc, br).t()
    opA = Am.t() if ta else Am
    opB = Bm.t() if tb else Bm
    out = torch.empty(n, m, dtype=torch.float64, device="cuda")
    ref = timeit(lambda: torch.matmul(opB.t(), opA.t(), out=out))
    err = float((out.t() - Cm.view(n, m).t()).abs().max() / out.abs().max())
    nbytes = 8.0 * (ar * ac + br * bc)
    print(f"{'NT'[ta]}{'NT'[tb]} {m:6d} {n:6d} {k:6d}  ours {ours*1e3:8.2f} us ({nbytes/ours*1e-6:6.0f} GB/s, {name})  cuBLAS {ref*1e3:8.2f} us  ratio {ref/ours:5.2f}  rel diff {err:.1e}", flush=True)
