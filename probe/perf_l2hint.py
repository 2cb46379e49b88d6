"""L2 eviction hints of the DGEMM's TMA loads (dgemm.l2_hints): time without ncu, DRAM traffic under ncu (SHORT=1: one launch per setting)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt

short = bool(os.environ.get("SHORT"))
h = rt.Handle(stream=torch.cuda.current_stream())
for (ta, tb, m, n, k) in [(1, 0, 8192, 8192, 8192), (0, 0, 8192, 8192, 8192), (0, 0, 16384, 16384, 4096)]:
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    A = torch.zeros(ar * ac, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    B = torch.zeros(br * bc, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
    C = torch.empty(m * n, dtype=torch.float64, device="cuda")
    ref = None
    for hints in (0, 1):
        h.set_option("dgemm.l2_hints", hints)
        run = lambda: h.gemm(True, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, C, m)
        run(); torch.cuda.synchronize()
        if ref is None:
            ref = C.clone()
        else:
            assert torch.equal(ref, C), "hints must not change the result"
        if short:
            continue
        best = 1e30
        for _ in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(); run(); b.record(); b.synchronize()
            best = min(best, a.elapsed_time(b) / 2)
        print(f"{'TN'[1-ta]}{'TN'[1-tb]} {m}x{n}x{k} l2_hints={hints}: {best:8.3f} ms {2.0*m*n*k/best*1e-9:6.2f} TF {h.last_kernel()}", flush=True)
    del A, B, C, ref
