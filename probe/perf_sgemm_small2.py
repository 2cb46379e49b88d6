"""Small SGEMMs: pre-split of op(B) on / off, CTA pairs on / off (fixed cost of the tcgen05 path), against cuBLAS FP32."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
from perf_tiles3 import timeit, h

torch.backends.cuda.matmul.allow_tf32 = False
for (m, n, k) in [(128, 128, 128), (256, 256, 256), (512, 512, 512), (768, 768, 768), (1024, 1024, 1024), (2048, 2048, 512), (512, 512, 4096), (2048, 256, 2048)]:
    A = torch.empty(m * k, dtype=torch.float32, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(k * n, dtype=torch.float32, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.zeros(m * n, dtype=torch.float32, device="cuda")
    res = {}
    for pre in (1, 0):
        for pair in (1, 0):
            h.set_option("sgemm.presplit", pre)
            h.set_option("sgemm.pair", pair)
            res[(pre, pair)] = timeit(lambda: h.gemm(False, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, Cm, m))
    h.set_option("sgemm.presplit", 1)
    h.set_option("sgemm.pair", 1)
    out = torch.empty(n, m, dtype=torch.float32, device="cuda")
    ref = timeit(lambda: torch.matmul(B.view(n, k), A.view(k, m), out=out))
    print(f"{m}x{n}x{k}: presplit+pair {res[(1,1)]*1e3:7.2f}  presplit {res[(1,0)]*1e3:7.2f}  pair {res[(0,1)]*1e3:7.2f}  neither {res[(0,0)]*1e3:7.2f} us   cuBLAS {ref*1e3:7.2f} us", flush=True)
