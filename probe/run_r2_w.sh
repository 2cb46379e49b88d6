#!/bin/bash
out=gpurun_out/r2w; mkdir -p $out
( time timeout 900 python -m pytest tests/test_gpu_l3.py tests/test_gpu_golden.py tests/test_gpu_parity.py -q -m gpu -k "syrk or sgemm or l3 or golden or trsm" ) > $out/pytest_l3.txt 2>&1; tail -5 $out/pytest_l3.txt
timeout 600 python probe/perf_l3.py 2>&1 | tee $out/perf_l3.txt
timeout 120 python - <<'PY' 2>&1 | tee $out/sgemm_big.txt
import sys, os
sys.path.insert(0, os.getcwd())
import torch, rtatblas_b200 as rt
h = rt.Handle(stream=torch.cuda.current_stream())
for n in (8192, 12288, 16384):
    A = torch.empty(n * n, dtype=torch.float32, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(n * n, dtype=torch.float32, device="cuda"); h.fill_uniform(B, 2)
    C = torch.empty(n * n, dtype=torch.float32, device="cuda")
    for _ in range(2): h.gemm(False, 0, 0, n, n, n, 1.0, A, n, B, n, 0.0, C, n)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): h.gemm(False, 0, 0, n, n, n, 1.0, A, n, B, n, 0.0, C, n)
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 3
    ref = (A.view(n, n)[:64].t().double() if False else None)
    print(f"sgemm NN {n}^3 {h.last_kernel()} {ms:.3f} ms {2.0*n**3/ms*1e-9:.1f} TFLOP/s", flush=True)
    del A, B, C
PY
echo done
