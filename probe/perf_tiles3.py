"""Tile configurations of the persistent DGEMM side by side (64x64, 128x64, 128x128; k-tail skipping on / off) on the shapes
where granularity decides, incl. config 3 as the planner's pad plan runs it (padded leading dimensions, operands hot in L2).
Each configuration's result is compared with the 64x64 one (max relative difference)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt

h = rt.Handle(stream=torch.cuda.current_stream())


def timeit(fn, reps=7, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); b.synchronize()
    batch = max(1, min(50, int(2.0 / max(a.elapsed_time(b), 1e-3))))
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(batch):
            fn()
        b.record(); b.synchronize()
        best = min(best, a.elapsed_time(b) / batch)
    return best


def case(name, ta, tb, m, n, k, lda=None, ldb=None, tiles=(64, 192, 128), ktail=(1,)):
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    lda, ldb = lda or ar, ldb or br
    A = torch.empty(lda * ac, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 1)
    B = torch.empty(ldb * bc, dtype=torch.float64, device="cuda"); h.fill_uniform(B, 2)
    Cm = torch.zeros(m * n, dtype=torch.float64, device="cuda")
    ref = None
    for t in tiles:
        for kt in ktail:
            h.set_option("dgemm.tile", t)
            h.set_option("dgemm.ktail", kt)
            Cm.zero_()
            fn = lambda: h.gemm(True, ta, tb, m, n, k, 1.0, A, lda, B, ldb, 0.0, Cm, m)
            ms = timeit(fn)
            torch.cuda.synchronize()
            if ref is None:
                ref = Cm.clone()
                diff = 0.0
            else:
                diff = float(((Cm - ref).abs().max() / ref.abs().max()).item())
            print(f"{name:40s} tile {t:3d} ktail {kt} {h.last_kernel():44s} {ms:9.4f} ms {2.0*m*n*k/ms*1e-9:7.2f} TF  maxdiff {diff:.1e}", flush=True)
    h.set_option("dgemm.tile", 0)
    h.set_option("dgemm.ktail", 1)


if __name__ == "__main__":
    case("C1 NN 1024^3", 0, 0, 1024, 1024, 1024)
    case("C3 pad plan TT-like 4097x3001x2049 (lda 2080, ldb 3008)", 1, 1, 4097, 3001, 2049, lda=2080, ldb=3008, ktail=(1, 0))
    case("C3 NT padded ld (4128, 3008)", 0, 1, 4097, 3001, 2049, lda=4128, ldb=3008, ktail=(1, 0))
    case("C3 NT unpadded (cp.async)", 0, 1, 4097, 3001, 2049, ktail=(1, 0))
    case("NN 2048^3", 0, 0, 2048, 2048, 2048)
    case("TN 2048^3", 1, 0, 2048, 2048, 2048)
    case("NN 4096^3", 0, 0, 4096, 4096, 4096)
    case("TN 8192^3", 1, 0, 8192, 8192, 8192)
    case("NN 8192x1024x8192", 0, 0, 8192, 1024, 8192)
    case("NN 8192x5120x1024 (e2e panel)", 0, 0, 8192, 5120, 1024)
    case("NN 512^3", 0, 0, 512, 512, 512)
    case("NN 3000x3000x3000 (ld 3008)", 0, 0, 3000, 3000, 3000, lda=3008, ldb=3008)
    case("NT 1000x1000x1000 unaligned", 0, 1, 1000, 1000, 1000)
