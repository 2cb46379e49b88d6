"""Config 4's shape (131072 x 128 x 4096) under the four op pairs, pair / single kernel: does A's stored orientation matter?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
h = rt.Handle(stream=torch.cuda.current_stream())

def timeit(fn, reps=5):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): fn()
        b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b) / 10)
    return best

for (m, n, k) in [(131072, 128, 4096), (65536, 256, 4096), (4096, 4096, 4096)]:
    for ta in (0, 1):
        for tb in (0, 1):
            ar, ac = (k, m) if ta else (m, k)
            br, bc = (n, k) if tb else (k, n)
            A = torch.empty(ar * ac, dtype=torch.float32, device="cuda"); h.fill_uniform(A, 1)
            B = torch.empty(br * bc, dtype=torch.float32, device="cuda"); h.fill_uniform(B, 2)
            C = torch.empty(m * n, dtype=torch.float32, device="cuda")
            for pair in (1, 0):
                h.set_option("sgemm.pair", pair)
                ms = timeit(lambda: h.gemm(False, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, C, m))
                print(f"{m}x{n}x{k} op {'T' if ta else 'N'}{'T' if tb else 'N'} pair={pair} {h.last_kernel():48s} {ms:8.4f} ms {2.0*m*n*k/ms*1e-9:8.1f} TFLOP/s", flush=True)
            del A, B, C
