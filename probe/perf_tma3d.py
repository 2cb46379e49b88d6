"""One 3-D TMA box per stage for M/N-contiguous DGEMM operands (dgemm.tma3d) against BMN / 16 two-dimensional boxes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt

h = rt.Handle(stream=torch.cuda.current_stream())
for (ta, tb, m, n, k) in [(0, 0, 8192, 8192, 8192), (0, 1, 8192, 8192, 8192), (1, 0, 8192, 8192, 8192), (0, 0, 1024, 1024, 1024),
                          (0, 0, 8192, 1024, 8192), (0, 1, 4112, 3008, 2049), (0, 0, 4096, 4096, 4096)]:
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    A = torch.zeros(ar * ac, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    B = torch.zeros(br * bc, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
    C = torch.empty(m * n, dtype=torch.float64, device="cuda")
    ref = None
    for mode in (0, 1):
        h.set_option("dgemm.tma3d", mode)
        run = lambda: h.gemm(True, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, C, m)
        C.fill_(float("nan")); run(); torch.cuda.synchronize()
        if ref is None:
            ref = C.clone()
        else:
            assert torch.equal(ref, C), "3-D boxes must not change the result"
        reps = 2 if m * n * k > 1e11 else 20
        best = 1e30
        for _ in range(5):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(reps):
                run()
            b.record(); b.synchronize()
            best = min(best, a.elapsed_time(b) / reps)
        print(f"{'TN'[1-ta]}{'TN'[1-tb]} {m}x{n}x{k} tma3d={mode}: {best:8.4f} ms {2.0*m*n*k/best*1e-9:6.2f} TF {h.last_kernel()}", flush=True)
    del A, B, C, ref
