"""ncu target for the SYRK / TRSM kernels: one DSYRK (n = k = 4096), one SSYRK, one DTRSM and one STRSM (4096 x 4096)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt

h = rt.Handle(stream=torch.cuda.current_stream())
n = 4096
for dt in (torch.float64, torch.float32):
    dbl = dt == torch.float64
    A = torch.empty(n * n, dtype=dt, device="cuda"); h.fill_uniform(A, 1)
    Cm = torch.zeros(n * n, dtype=dt, device="cuda")
    h.syrk(dbl, 0, 0, n, n, 1.0, A, n, 0.0, Cm, n)
    T = A.clone(); T.view(n, n).mul_(1.0 / 64).diagonal().add_(2.0)
    B = torch.empty(n * n, dtype=dt, device="cuda"); h.fill_uniform(B, 3)
    h.trsm(dbl, 0, 0, 0, 0, n, n, 1.0, T, n, B, n)
torch.cuda.synchronize()
print("ok")
