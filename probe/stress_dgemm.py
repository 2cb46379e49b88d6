"""Iteration-count stress of the DGEMM kernels: the same call repeated thousands of times, every result bit-compared with
the first (the kernels are deterministic: stream-K partials are summed in CTA order), optionally with copy-engine traffic
on other streams (what the host pipeline adds).  Reports every mismatch with the rows / columns it covers.

usage: stress_dgemm.py [--iters N] [--traffic 0|1] [--tile 0|64|128] [--pipeline REPS] [--cases a,b,..]
RTAT_B200_LIB selects an experiment build of the library."""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import rtatblas_b200 as rt

ap = argparse.ArgumentParser()
ap.add_argument("--iters", type=int, default=2000)
ap.add_argument("--traffic", type=int, default=1)
ap.add_argument("--tile", type=int, default=0)
ap.add_argument("--pipeline", type=int, default=0)
ap.add_argument("--cases", default="")
args = ap.parse_args()

torch.cuda.set_device(0)
handle = rt.Handle(stream=torch.cuda.current_stream())
if args.tile:
    handle.set_option("dgemm.tile", args.tile)

# (name, ta, tb, m, n, k, lda_pad, ldc_pad, alpha, beta)
CASES = [
    ("tn_1000x256x3000", 1, 0, 1000, 256, 3000, 0, 5, 1.25, -0.5),   # the scenario of DESIGN r1 §8.1
    ("tn_1000x191x3000", 1, 0, 1000, 191, 3000, 0, 5, 1.25, -0.5),
    ("nn_1024", 0, 0, 1024, 1024, 1024, 0, 0, 1.0, 0.0),
    ("nt_777x513x2049_odd", 0, 1, 777, 513, 2049, 0, 1, 1.0, 0.0),     # odd ld -> cp.async producers
    ("tt_640x384x1536", 1, 1, 640, 384, 1536, 2, 0, 1.0, 1.0),
    ("nn_2048x2048x512", 0, 0, 2048, 2048, 512, 0, 0, 1.0, 0.0),
    ("tn_4096x1024x1024", 1, 0, 4096, 1024, 1024, 0, 0, 1.0, 1.0),
    # pattern: A(i, k) = k // 16 (the k-tile index), B = 1, so a fragment taken from another k-tile shows as a signed
    # multiple of the tile distance: negative = stale (an older fill of the stage), positive = a newer fill
    ("pattern_tn_4096x1024x1024", 1, 0, 4096, 1024, 1024, 0, 0, 1.0, 0.0),
    ("pattern_nn_2048x2048x512", 0, 0, 2048, 2048, 512, 0, 0, 1.0, 0.0),
]
if args.cases:
    want = set(args.cases.split(","))
    CASES = [c for c in CASES if c[0] in want]

traffic = None
if args.traffic:
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    h_src = torch.empty(4 << 20, dtype=torch.float64).pin_memory()
    h_dst = torch.empty(4 << 20, dtype=torch.float64).pin_memory()
    d_a = torch.empty(4 << 20, dtype=torch.float64, device="cuda")
    d_b = torch.zeros(4 << 20, dtype=torch.float64, device="cuda")

    def traffic():
        with torch.cuda.stream(s_in):
            d_a.copy_(h_src, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_dst.copy_(d_b, non_blocking=True)

total_bad = 0
for name, ta, tb, m, n, k, lda_pad, ldc_pad, alpha, beta in CASES:
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    lda, ldb, ldc = ar + lda_pad, br, m + ldc_pad
    g = torch.Generator(device="cuda").manual_seed(7)
    A = torch.rand(lda * ac, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    B = torch.rand(ldb * bc, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    C0 = torch.rand(ldc * n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    pattern = name.startswith("pattern")
    if pattern:
        kt = (torch.arange(k, device="cuda") // 16).to(torch.float64)
        A = (kt.repeat(m) if ta else kt.repeat_interleave(m)).contiguous()   # stored k x m (T) or m x k (N), tight ld
        B = torch.ones(ldb * bc, dtype=torch.float64, device="cuda")
    C = C0.clone()
    handle.gemm(True, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
    kern = handle.last_kernel()
    ref = C.clone()
    if pattern:  # the exact answer, so a corrupted first run cannot hide: every entry is 16 * sum of the k-tile indices
        T = k // 16
        ref = torch.full_like(C, 16.0 * T * (T - 1) / 2)
    torch.cuda.synchronize()
    nbad = torch.zeros((), dtype=torch.int64, device="cuda")
    first_bad = None
    t0 = time.time()
    for it in range(args.iters):
        if traffic and it % 4 == 0:
            traffic()
        C.copy_(C0)
        handle.gemm(True, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
        neq = C != ref
        nbad += neq.any()
        if it % 250 == 249 or it == args.iters - 1:
            if int(nbad.item()) and first_bad is None:
                # localise in the most recent batch: rerun until it shows (bounded)
                first_bad = it
    torch.cuda.synchronize()
    dt = time.time() - t0
    bad = int(nbad.item())
    total_bad += bad
    print(f"stress {name:24s} {kern:44s} iters {args.iters} traffic {args.traffic} mismatching runs {bad}  ({dt:.1f} s)", flush=True)
    shown = 0
    if bad:
        # characterise: repeat with a host check per iteration until one shows
        for it in range(4 * args.iters):
            if traffic and it % 4 == 0:
                traffic()
            C.copy_(C0)
            handle.gemm(True, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
            neq = (C != ref).view(n, ldc)[:, :m]
            if bool(neq.any().item()):
                idx = neq.nonzero()
                cols, rows = sorted(set(idx[:, 0].tolist())), sorted(set(idx[:, 1].tolist()))
                d = (C - ref).view(n, ldc)[:, :m][neq]
                print(f"   iteration {it}: {idx.shape[0]} elements, rows {rows[:12]}{'...' if len(rows) > 12 else ''} "
                      f"cols {cols[0]}..{cols[-1]} ({len(cols)} distinct), |delta| {d.abs().min().item():.3e}..{d.abs().max().item():.3e}", flush=True)
                if pattern:
                    vals, counts = torch.unique(d, return_counts=True)
                    print("   pattern deltas (value: count): " + ", ".join(f"{v:g}: {c}" for v, c in zip(vals.tolist()[:12], counts.tolist()[:12])), flush=True)
                    shown += 1
                    if shown < 6:
                        continue
                break

# the host pipeline itself (multi-GPU entry at world 1, as probe/debug_mg1.py): same inputs every repetition, bit-compare
if args.pipeline:
    cs = rt.ColumnSplit(handle, 0, 1)
    ta, tb, m, n, k = 1, 0, 1000, 703, 3000
    cs.host_setup(k * m * 8)
    cs.host_connect(None)
    g = torch.Generator().manual_seed(11)
    hA = torch.rand(k * m, dtype=torch.float64, generator=g).mul_(2).sub_(1).pin_memory()
    hB = torch.rand(n * k, dtype=torch.float64, generator=g).mul_(2).sub_(1).pin_memory()
    ldc = m + 5
    hC0 = torch.rand(ldc * n, dtype=torch.float64, generator=g).pin_memory()
    hC = hC0.clone().pin_memory()
    ref = None
    bad = 0
    t0 = time.time()
    for rep in range(args.pipeline):
        hC.copy_(hC0)
        cs.dgemm_host(ta, tb, m, n, k, 1.25, hA, k, hB, k, -0.5, hC, ldc)
        if ref is None:
            ref = hC.clone()
            continue
        neq = (hC != ref).view(n, ldc)[:, :m]
        if bool(neq.any()):
            bad += 1
            idx = neq.nonzero()
            cols, rows = sorted(set(idx[:, 0].tolist())), sorted(set(idx[:, 1].tolist()))
            if bad <= 5:
                print(f"   pipeline rep {rep}: {idx.shape[0]} elements, rows {rows[:12]} cols {cols[0]}..{cols[-1]} ({len(cols)} distinct)", flush=True)
    total_bad += bad
    print(f"stress pipeline tn_1000x703x3000 reps {args.pipeline} mismatching runs {bad} ({time.time() - t0:.1f} s) last kernel {handle.last_kernel()}", flush=True)

print(f"STRESS TOTAL mismatching runs: {total_bad}", flush=True)
sys.exit(1 if total_bad else 0)
