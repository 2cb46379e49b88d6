"""e2e schedule sweep for the host entry point on config 2 (8192^3 TN, pinned host buffers)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt

h = rt.Handle(stream=torch.cuda.current_stream())
n = 8192
hA = torch.rand(n * n, dtype=torch.float64).pin_memory()
hB = torch.rand(n * n, dtype=torch.float64).pin_memory()
hC = torch.zeros(n * n, dtype=torch.float64).pin_memory()


def run(reps=3):
    h.gemm(True, 1, 0, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n, host=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        h.gemm(True, 1, 0, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n, host=True)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)


for panels in (0, 2, 4, 8):
    for lead in (0, 2, 3, 4, 5, 6):
        h.set_option("host.panels", panels)
        h.set_option("host.lead", lead)
        ms = run()
        print(f"panels={panels} lead={lead}: {ms:.2f} ms  {2.0 * n**3 / ms / 1e9:.1f} TF", flush=True)
