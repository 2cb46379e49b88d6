"""ncu target: config 4 once through the single-CTA and once through the CTA-pair SGEMM kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
h = rt.Handle(stream=torch.cuda.current_stream())
m, n, k = 131072, 128, 4096
A = torch.empty(m * k, dtype=torch.float32, device="cuda"); h.fill_uniform(A, 1)
B = torch.empty(k * n, dtype=torch.float32, device="cuda"); h.fill_uniform(B, 2)
C = torch.empty(m * n, dtype=torch.float32, device="cuda")
for pair in (0, 1, 0, 1):
    h.set_option("sgemm.pair", pair)
    h.gemm(False, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, C, m)
torch.cuda.synchronize()
print("ok", h.last_kernel())
