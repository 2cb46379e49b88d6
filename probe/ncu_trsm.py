import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
h = rt.Handle(stream=torch.cuda.current_stream())
m, n = 128, 8192
A = torch.empty(m * m, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 2)
A.view(m, m).mul_(1.0 / 64).diagonal().add_(2.0)
B = torch.empty(m * n, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 3)
for _ in range(3):
    h.trsm(True, 0, 0, 0, 0, m, n, 1.0, A, m, B, m)
torch.cuda.synchronize()
print("ok", h.last_kernel())
