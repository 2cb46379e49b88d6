"""ncu workload of round 2's last kernels: config 3 as the planner's pad plan runs it (128x64 tiles + skinny strip),
4096^3 on 128x64 tiles, config 1 on 64x64 tiles, and a GEMV-shaped call on either skinny kernel."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
h = rt.Handle(stream=torch.cuda.current_stream())


def run(ta, tb, m, n, k, lda=None, ldb=None, reps=1):
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    lda, ldb = lda or ar, ldb or br
    A = torch.empty(lda * ac, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(ldb * bc, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.empty(m * n, dtype=torch.float64, device="cuda")
    for _ in range(reps):
        h.gemm(True, ta, tb, m, n, k, 1.0, A, lda, B, ldb, 0.0, Cm, m)
    torch.cuda.synchronize()
    print(h.last_kernel())


run(1, 1, 4097, 3001, 2049, lda=2080, ldb=3008)   # config 3, plan TNNPPN / NTNPPN: the multiply
run(0, 0, 4096, 4096, 4096)                        # 128x64 tiles, 7 waves
run(0, 0, 1024, 1024, 1024, reps=2)                # config 1
run(0, 1, 1, 8192, 4096)                           # GEMV, X contiguous along N
run(0, 0, 1, 8192, 2048)                           # GEMV, X contiguous along K
