"""SGEMM 131072 x n x 4096 for n below the 128-wide tile: the MMA work per tile stays that of n = 128 (zero-filled
columns) while the bytes TMA pulls for B shrink with n — separates tensor-bound from data-delivery-bound time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
h = rt.Handle(stream=torch.cuda.current_stream())
m, k = 131072, 4096
A = torch.empty(m * k, dtype=torch.float32, device="cuda"); h.fill_uniform(A, 1)
for n in (32, 64, 96, 128, 256):
    B = torch.empty(n * k, dtype=torch.float32, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.empty(m * n, dtype=torch.float32, device="cuda")
    for _ in range(2):
        h.gemm(False, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, Cm, m)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        h.gemm(False, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, Cm, m)
    b.record(); torch.cuda.synchronize()
    print(f"n={n}: {a.elapsed_time(b) / 5:.3f} ms  [{h.last_kernel()}]", flush=True)
