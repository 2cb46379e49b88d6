"""ncu target: config 3 (4097 x 3001 x 2049, op NT, odd leading dimensions) on the cp.async-producer DGEMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
h = rt.Handle(stream=torch.cuda.current_stream())
m, n, k = 4097, 3001, 2049
A = torch.empty(m * k, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
B = torch.empty(n * k, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
Cm = torch.empty(m * n, dtype=torch.float64, device="cuda")
tile = int(os.environ.get("TILE", "0"))
h.set_option("dgemm.tile", tile)
for _ in range(2):
    h.gemm(True, 0, 1, m, n, k, 1.0, A, m, B, n, 0.0, Cm, m)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5):
    h.gemm(True, 0, 1, m, n, k, 1.0, A, m, B, n, 0.0, Cm, m)
b.record(); torch.cuda.synchronize()
print("tile", tile, h.last_kernel(), a.elapsed_time(b) / 5, "ms")
