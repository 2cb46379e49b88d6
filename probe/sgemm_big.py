"""Large square SGEMMs (A no longer fits L2): what the tile raster costs in DRAM traffic."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, rtatblas_b200 as rt
h = rt.Handle(stream=torch.cuda.current_stream())
for (m, n, k, ta, tb) in [(8192, 8192, 8192, 0, 0), (12288, 12288, 12288, 0, 0), (16384, 16384, 16384, 0, 0), (16384, 16384, 16384, 1, 0), (32768, 4096, 8192, 0, 1)]:
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    A = torch.empty(ar * ac, dtype=torch.float32, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(br * bc, dtype=torch.float32, device="cuda"); h.fill_uniform(B, 2)
    C = torch.empty(m * n, dtype=torch.float32, device="cuda")
    for pair in (1, 0):
        h.set_option("sgemm.pair", pair)
        for _ in range(2): h.gemm(False, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, C, m)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): h.gemm(False, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, C, m)
        b.record(); b.synchronize()
        ms = a.elapsed_time(b) / 3
        print(f"sgemm {'T' if ta else 'N'}{'T' if tb else 'N'} {m}x{n}x{k} pair={pair} {h.last_kernel()} {ms:.3f} ms {2.0*m*n*k/ms*1e-9:.1f} TFLOP/s", flush=True)
    del A, B, C
