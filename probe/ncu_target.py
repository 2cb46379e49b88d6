"""Small fixed workload for ncu captures: C1 (1024^3 NN) x3, C2 (8192^3 TN) x1, C3 x1 through the C ABI."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
h = rt.Handle(stream=torch.cuda.current_stream())
which = sys.argv[1:] or ["c1", "c2", "c3", "c4", "geam"]
def run(ta, tb, m, n, k, reps):
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    A = torch.empty(ar * ac, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(br * bc, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.empty(m * n, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        h.gemm(True, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, Cm, m)
    torch.cuda.synchronize()
if "c1" in which: run(0, 0, 1024, 1024, 1024, 3)
if "c2" in which: run(1, 0, 8192, 8192, 8192, 1)
if "c3" in which: run(0, 1, 4097, 3001, 2049, 1)
if "c4" in which:
    m, n, k = 131072, 128, 4096
    A = torch.empty(m * k, dtype=torch.float32, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(k * n, dtype=torch.float32, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.empty(m * n, dtype=torch.float32, device="cuda")
    h.gemm(False, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, Cm, m)
    torch.cuda.synchronize()
if "geam" in which:
    n = 8192
    A = torch.empty(n * n, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    Cm = torch.empty(n * n, dtype=torch.float64, device="cuda")
    h.geam(True, 1, 0, n, n, 1.0, A, n, 0.0, Cm, n, Cm, n)
    h.geam(True, 0, 0, n, n, 1.0, A, n, 0.0, Cm, n, Cm, n)
    torch.cuda.synchronize()
print("done")
