"""Tile-choice heuristic check: automatic choice against each forced configuration (64x64, 128x64 = 192, 128x128)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
from perf_tiles3 import timeit, h

SHAPES = [
    (0, 0, 1024, 1024, 1024), (0, 0, 1536, 1536, 1536), (0, 0, 2048, 2048, 2048), (1, 0, 3072, 3072, 3072), (0, 0, 4096, 4096, 4096),
    (0, 1, 6144, 6144, 6144), (1, 0, 8192, 8192, 8192), (0, 0, 1000, 1000, 1000), (0, 0, 3000, 3000, 3000), (0, 1, 5000, 5000, 5000),
    (0, 0, 131072, 128, 4096), (0, 0, 128, 131072, 4096), (0, 0, 1024, 8192, 1024), (0, 0, 8192, 1024, 1024), (0, 0, 4096, 4096, 512),
    (1, 0, 8192, 5120, 256), (1, 0, 8192, 5120, 512), (1, 0, 8192, 5120, 1024), (0, 0, 32768, 4096, 2048), (0, 0, 32768, 1536, 512),
    (0, 0, 2048, 2048, 8192), (0, 0, 512, 512, 16384), (0, 1, 4097, 3001, 2049), (1, 1, 2500, 7000, 1500), (0, 0, 8192, 8192, 512),
]
worst = 1.0
for (ta, tb, m, n, k) in SHAPES:
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    lda, ldb = (ar + 31) // 32 * 32, (br + 31) // 32 * 32
    A = torch.empty(lda * ac, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(ldb * bc, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.zeros(m * n, dtype=torch.float64, device="cuda")
    res = {}
    for t in (0, 64, 192, 128):
        h.set_option("dgemm.tile", t)
        ms = timeit(lambda: h.gemm(True, ta, tb, m, n, k, 1.0, A, lda, B, ldb, 0.0, Cm, m), reps=5)
        res[t] = (ms, h.last_kernel())
    h.set_option("dgemm.tile", 0)
    best = min(res[t][0] for t in (64, 192, 128))
    worst = max(worst, res[0][0] / best)
    pick = res[0][1].replace("dgemm_ws_dmma_", "")
    print(f"{'NT'[ta]}{'NT'[tb]} {m:6d} {n:6d} {k:6d}  auto {res[0][0]:9.4f} ms ({pick:32s})  64: {res[64][0]:9.4f}  128x64: {res[192][0]:9.4f}  128: {res[128][0]:9.4f}"
          f"  auto/best {res[0][0] / best:.3f}  {2.0*m*n*k/res[0][0]*1e-9:6.2f} TF", flush=True)
    del A, B, Cm
print("worst auto / best:", round(worst, 3))
