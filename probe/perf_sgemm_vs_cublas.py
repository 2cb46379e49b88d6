"""SGEMM (3xTF32 tcgen05) against cuBLAS FP32 (torch.matmul, TF32 off) over sizes: where does the reference's path win?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
from perf_tiles3 import timeit, h

torch.backends.cuda.matmul.allow_tf32 = False
for (ta, tb, m, n, k) in [(0, 0, 256, 256, 256), (0, 0, 512, 512, 512), (0, 0, 1024, 1024, 1024), (1, 0, 2048, 2048, 2048), (0, 1, 4096, 4096, 4096),
                          (0, 0, 1000, 1000, 1000), (0, 0, 4097, 3001, 2049), (0, 0, 131072, 128, 4096), (0, 0, 128, 131072, 4096), (0, 0, 8192, 64, 8192),
                          (0, 0, 1, 8192, 4096), (0, 0, 8192, 8192, 256)]:
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    A = torch.empty(ar * ac, dtype=torch.float32, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(br * bc, dtype=torch.float32, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.zeros(m * n, dtype=torch.float32, device="cuda")
    ours = timeit(lambda: h.gemm(False, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, Cm, m))
    name = h.last_kernel()
    Am = A.view(ac, ar).t(); Bm = B.view(bc, br).t()
    opA = Am.t() if ta else Am
    opB = Bm.t() if tb else Bm
    out = torch.empty(n, m, dtype=torch.float32, device="cuda")
    ref = timeit(lambda: torch.matmul(opB.t(), opA.t(), out=out))
    err = float((out.t() - Cm.view(n, m).t()).abs().max() / out.abs().max())
    print(f"{'NT'[ta]}{'NT'[tb]} {m:6d} {n:6d} {k:6d}  ours {ours*1e3:9.2f} us {2.0*m*n*k/ours*1e-9:7.1f} TF ({name})  cuBLAS {ref*1e3:9.2f} us  ratio {ref/ours:5.2f}  rel diff {err:.1e}", flush=True)
