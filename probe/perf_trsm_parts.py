"""Where TRSM time goes: single block solves and the small-k GEMM updates the recursion issues."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import rtatblas_b200 as rt

h = rt.Handle(stream=torch.cuda.current_stream())


def timeit(fn, reps=20):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


for dt in (torch.float64, torch.float32):
    dbl = dt == torch.float64
    print("dtype", dt)
    for m, n, left in [(32, 8192, 1), (128, 8192, 1), (128, 1024, 1), (256, 8192, 1), (8192, 128, 0), (1024, 128, 0)]:
        na = m if left else n
        A = torch.empty(na * na, dtype=dt, device="cuda")
        h.fill_uniform(A, 2)
        A.view(na, na).mul_(1.0 / 64).diagonal().add_(2.0)
        B = torch.empty(m * n, dtype=dt, device="cuda")
        h.fill_uniform(B, 3)
        us = timeit(lambda: h.trsm(dbl, 0 if left else 1, 0, 0, 0, m, n, 1.0, A, na, B, m))
        print(f"  trsm m={m} n={n} left={left}: {us:.1f} us")
    for m, n, k in [(128, 8192, 128), (256, 8192, 256), (512, 8192, 512), (1024, 8192, 1024), (2048, 8192, 2048), (4096, 8192, 4096), (128, 1024, 128), (512, 1024, 512)]:
        A = torch.empty(m * k, dtype=dt, device="cuda")
        B = torch.empty(k * n, dtype=dt, device="cuda")
        Cm = torch.empty(m * n, dtype=dt, device="cuda")
        h.fill_uniform(A, 1); h.fill_uniform(B, 2); h.fill_uniform(Cm, 3)
        us = timeit(lambda: h.gemm(dbl, 0, 0, m, n, k, -1.0, A, m, B, k, 1.0, Cm, m))
        print(f"  gemm {m}x{n}x{k} beta=1: {us:.1f} us  {2.0 * m * n * k / us / 1e6:.1f} TF  [{h.last_kernel()}]")
