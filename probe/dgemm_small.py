"""DGEMM kernel time on small / mid cubes (CUDA events, 20 back-to-back launches)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
if os.environ.get('RTAT_VARIANT_LIB'):
    import rtatblas_b200._capi as _c
    _c.lib_path = lambda: os.environ['RTAT_VARIANT_LIB']
h = rt.Handle(stream=torch.cuda.current_stream())
for (m, n, k, ta, tb) in [(1024, 1024, 1024, 0, 0), (1024, 1024, 1024, 1, 0), (2048, 2048, 2048, 0, 0), (4096, 4096, 4096, 1, 0), (8192, 8192, 8192, 1, 0)]:
    A = torch.empty(m * k, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(n * k, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.empty(m * n, dtype=torch.float64, device="cuda")
    f = lambda: h.gemm(True, ta, tb, m, n, k, 1.0, A, k if ta else m, B, n if tb else k, 0.0, Cm, m)
    f(); f(); torch.cuda.synchronize()
    reps = 20 if m <= 2048 else 3
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    print(f"{m}x{n}x{k} ta={ta} tb={tb}: {ms*1e3:.1f} us  {2.0*m*n*k/ms/1e9:.2f} TF  [{h.last_kernel()}]", flush=True)
