"""e2e schedule sweep (round 2, session 3): growing K-panel head forced on (host.panels = -g) x leading blocks, config 2."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt

h = rt.Handle(stream=torch.cuda.current_stream())
n = 8192
hA = torch.rand(n * n, dtype=torch.float64).pin_memory()
hB = torch.rand(n * n, dtype=torch.float64).pin_memory()
hC = torch.zeros(n * n, dtype=torch.float64).pin_memory()


def run(reps=4):
    h.gemm(True, 1, 0, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n, host=True)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        h.gemm(True, 1, 0, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n, host=True)
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts), sorted(ts)[len(ts) // 2]


for panels, lead in [(0, 0), (0, 5), (-1, 4), (-10, 4), (-1, 5), (-15, 5), (-25, 5), (-1, 6), (-25, 6), (-40, 6), (-25, 7), (0, 0)]:
    h.set_option("host.panels", panels)
    h.set_option("host.lead", lead)
    sc = rt.host_schedule(8, 1, n, n, n, n, 0, panels, lead)
    ms, med = run()
    print(f"panels={panels} lead={lead}: min {ms:.2f} ms median {med:.2f}  {2.0 * n**3 / ms / 1e9:.1f} TF  npanels {len(sc['panels'])} first {sc['panels'][:3]}", flush=True)
h.set_option("host.trace", 1)
for panels, lead in [(0, 0), (-15, 5)]:
    h.set_option("host.panels", panels)
    h.set_option("host.lead", lead)
    print(f"--- trace panels={panels} lead={lead}", file=sys.stderr, flush=True)
    h.gemm(True, 1, 0, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n, host=True)
    torch.cuda.synchronize()
