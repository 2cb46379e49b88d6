"""ncu target: config 4 (SGEMM 131072 x 128 x 4096, op NN) on the tcgen05 3xTF32 kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
if os.environ.get('RTAT_VARIANT_LIB'):
    import rtatblas_b200._capi as _c
    _c.lib_path = lambda: os.environ['RTAT_VARIANT_LIB']
h = rt.Handle(stream=torch.cuda.current_stream())
m, n, k = 131072, 128, 4096
A = torch.empty(m * k, dtype=torch.float32, device="cuda"); h.fill_uniform(A, 1)
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
print(h.last_kernel(), a.elapsed_time(b) / 5, "ms")
