"""NN against TN at 8192^3 for an ncu capture: A M-contiguous vs K-contiguous."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
h = rt.Handle(stream=torch.cuda.current_stream())
n = 8192
A = torch.zeros(n * n, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
B = torch.zeros(n * n, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
C = torch.empty(n * n, dtype=torch.float64, device="cuda")
for ta, tb in ((0, 0), (1, 0), (0, 1)):
    h.gemm(True, ta, tb, n, n, n, 1.0, A, n, B, n, 0.0, C, n)
torch.cuda.synchronize()
print("done")
