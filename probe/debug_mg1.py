import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import rtatblas_b200 as rt
torch.cuda.set_device(0)
handle = rt.Handle(stream=torch.cuda.current_stream())
cs = rt.ColumnSplit(handle, 0, 1)
which = ["0", "1"]
if len(sys.argv) > 1: handle.set_option("dgemm.tile", int(sys.argv[1]))
cases = [(0, 0, 4096, 1024, 4096), (1, 0, 1000, 703, 3000)]
hb = cs.host_setup(max((k if ta else m) * (m if ta else k) * 8 for ta, _, m, _, k in cases)); cs.host_connect(None)
for ci, (ta, tb, m, n, k) in enumerate(cases):
    if str(ci) not in which: continue
    ar, ac = (k, m) if ta else (m, k)
    for rep in range(6):
        g = torch.Generator().manual_seed(1000 * ci + rep)
        hA = torch.rand(ar * ac, dtype=torch.float64, generator=g).mul_(2).sub_(1).pin_memory()
        hB = torch.rand(n * k, dtype=torch.float64, generator=g).mul_(2).sub_(1).pin_memory()
        ldb, ldc = k, m + 5
        hC = torch.full((ldc * n,), 3.0, dtype=torch.float64).pin_memory()
        beta = 0.0 if rep == 0 else -0.5
        cs.dgemm_host(ta, tb, m, n, k, 1.25, hA, ar, hB, ldb, beta, hC, ldc)
        dA, dB = hA.cuda(), hB.cuda()
        opA = dA.view(ac, ar) if not ta else dA.view(ac, ar).t()
        opB = dB.view(n, ldb)
        want = 1.25 * (opB @ opA) + beta * 3.0
        got = hC.view(n, ldc)[:, :m].cuda()
        err = (got - want).abs()
        bad = (err > 1e-8).nonzero()
        msg = "ok"
        if bad.shape[0]:
            msg = f"BAD {bad.shape[0]}: cols(j) {bad[:,0].min().item()}..{bad[:,0].max().item()} rows(i) {bad[:,1].min().item()}..{bad[:,1].max().item()}; sample got {got[bad[0,0], bad[0,1]].item():.4f} want {want[bad[0,0], bad[0,1]].item():.4f}"
            js = sorted(set(bad[:,0].tolist())); is_ = sorted(set(bad[:,1].tolist()))
            msg += f" | distinct cols {len(js)} first {js[:6]} | distinct rows {len(is_)} first {is_[:10]}"
        print(f"case {ci} rep {rep} beta {beta}: max err {err.max().item():.2e} {handle.last_kernel()} {msg}", flush=True)
