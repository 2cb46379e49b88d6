"""Pad copies of config 3 (MatrixMove into a padded leading dimension): streaming geam, equal chunks on / off, cold and warm."""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import rtatblas_b200 as rt

h = rt.Handle(stream=torch.cuda.current_stream())
flush_buf = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")


def timed(fn, reps, flush):
    out = []
    for _ in range(reps):
        if flush:
            flush_buf.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); b.synchronize()
        out.append(a.elapsed_time(b))
    return statistics.median(out)


for (rows, cols, ldo) in [(4097, 2049, 4128), (3001, 2049, 3008), (4097, 3001, 4128), (8191, 8192, 8192), (1000, 1000, 1024), (4096, 2049, 4128)]:
    A = torch.empty(rows * cols, dtype=torch.float64, device="cuda"); h.fill_uniform(A, 3)
    out = torch.empty(ldo * cols, dtype=torch.float64, device="cuda")
    fn = lambda: h.geam(True, 0, 0, rows, cols, 1.0, A, rows, 0.0, out, ldo, out, ldo)
    nbytes = 2.0 * rows * cols * 8
    for ev in (0, 1, 2):
        h.set_option("geam.even_chunks", min(ev, 1))
        h.set_option("geam.persistent", 1 if ev == 2 else 0)
        for _ in range(3):
            fn()
        cold, warm = timed(fn, 30, True), timed(fn, 30, False)
        ok = bool(torch.equal(out.view(cols, ldo)[:, :rows], A.view(cols, rows)))
        print(f"{rows}x{cols} -> ld {ldo}  even_chunks/persistent {ev} {h.last_kernel():20s} cold {cold*1e3:7.2f} us {nbytes/cold*1e-6:7.0f} GB/s   warm {warm*1e3:7.2f} us {nbytes/warm*1e-6:7.0f} GB/s  exact {ok}", flush=True)
    h.set_option("geam.even_chunks", 1)
    h.set_option("geam.persistent", 0)
