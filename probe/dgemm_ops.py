"""DGEMM 8192^3 for the four op pairs (which operand is K-contiguous decides how many TMA boxes a stage needs)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
if os.environ.get('RTAT_VARIANT_LIB'):
    import rtatblas_b200._capi as _c
    _c.lib_path = lambda: os.environ['RTAT_VARIANT_LIB']
h = rt.Handle(stream=torch.cuda.current_stream())
n = int(os.environ.get("N", "8192"))
A = torch.empty(n * n, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
B = torch.empty(n * n, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
Cm = torch.empty(n * n, dtype=torch.float64, device="cuda")
for ta, tb in [(1, 0), (0, 0), (0, 1), (1, 1)]:
    f = lambda: h.gemm(True, ta, tb, n, n, n, 1.0, A, n, B, n, 0.0, Cm, n)
    f(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): f()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"op {'NT'[ta]}{'NT'[tb]}: {ms:.3f} ms  {2.0*n**3/ms/1e9:.2f} TF  [{h.last_kernel()}]", flush=True)
