"""Small SGEMMs: tcgen05 3xTF32 kernel (sgemm.variant 2) against the FP32 SIMT kernel (1) and cuBLAS FP32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
from perf_tiles3 import timeit, h

torch.backends.cuda.matmul.allow_tf32 = False
for n in (128, 256, 384, 512, 768, 1024):
    A = torch.empty(n * n, dtype=torch.float32, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(n * n, dtype=torch.float32, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.zeros(n * n, dtype=torch.float32, device="cuda")
    res = {}
    for v in (2, 1):
        h.set_option("sgemm.variant", v)
        res[v] = timeit(lambda: h.gemm(False, 0, 0, n, n, n, 1.0, A, n, B, n, 0.0, Cm, n))
    h.set_option("sgemm.variant", 0)
    out = torch.empty(n, n, dtype=torch.float32, device="cuda")
    ref = timeit(lambda: torch.matmul(B.view(n, n), A.view(n, n), out=out))
    print(f"{n}^3: tcgen05 {res[2]*1e3:7.2f} us  simt {res[1]*1e3:7.2f} us  cuBLAS {ref*1e3:7.2f} us", flush=True)
