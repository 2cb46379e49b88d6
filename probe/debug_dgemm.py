import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt
h = rt.Handle(stream=torch.cuda.current_stream())
def case(ta, tb, m, n, k, alpha, beta, ldc_pad, tile):
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    ldc = m + ldc_pad
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.rand(ar * ac, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    B = torch.rand(br * bc, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    C0 = torch.rand(ldc * n, dtype=torch.float64, device="cuda", generator=g)
    Cm = C0.clone()
    h.set_option("dgemm.tile", tile)
    h.gemm(True, ta, tb, m, n, k, alpha, A, ar, B, br, beta, Cm, ldc)
    torch.cuda.synchronize()
    Am, Bm = A.view(ac, ar).t(), B.view(bc, br).t()
    opA, opB = (Am.t() if ta else Am), (Bm.t() if tb else Bm)
    ref = alpha * (opA @ opB) + beta * C0.view(n, ldc).t()[:m]
    err = (Cm.view(n, ldc).t()[:m] - ref).abs()
    bad = (err > 1e-9).nonzero()
    msg = ""
    if bad.shape[0]:
        msg = f" BAD {bad.shape[0]} rows {bad[:,0].min().item()}..{bad[:,0].max().item()} cols {bad[:,1].min().item()}..{bad[:,1].max().item()}"
    print(f"{'TN'[1-ta]}{'TN'[1-tb]} {m}x{n}x{k} a={alpha} b={beta} ldc+{ldc_pad} tile {tile}: max err {err.max().item():.2e} {h.last_kernel()}{msg}", flush=True)
for tile in (0, 64, 128):
    case(1, 0, 1000, 256, 3000, 1.25, -0.5, 5, tile)
    case(1, 0, 1000, 191, 3000, 1.25, -0.5, 5, tile)
    case(1, 0, 1000, 256, 3000, 1.25, 0.0, 5, tile)
    case(1, 0, 1000, 256, 2992, 1.25, -0.5, 5, tile)
    case(1, 0, 1024, 256, 3000, 1.25, -0.5, 0, tile)
    case(0, 0, 1000, 256, 3000, 1.0, 0.0, 0, tile)
