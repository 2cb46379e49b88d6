"""-m gpu iteration-count stress: every shipped DGEMM tile configuration x producer kind x op pair runs the same call
2000 times and every result is bit-compared with the first (the kernels are deterministic: stream-K partials are added
in CTA order).  Parity cases run a kernel once or twice; this is what catches a race that needs a million k-tiles to show.

Background: a consumer warp used to release a shared-memory ring stage while its last fragment load of
that stage was still in flight, and with 16-byte fragment loads on 64x64 tiles (two CTAs per SM) 4-14 % of runs of
4096 x 1024 x 1024 took a fragment from the NEXT fill of the stage.  The shapes below keep that regime: more tiles than
two waves of CTAs (data-parallel tiles follow each CTA's stream-K run) and tens of ring wrap-arounds per segment.
The `pattern` cases use A(i, k) = k // 16 and B = 1, whose exact product is known and in which a fragment from a wrong
k-tile shows up as a multiple of the tile distance."""
import pytest
import torch

pytestmark = pytest.mark.gpu

ITERS = 2000
OPS = [(0, 0), (0, 1), (1, 0), (1, 1)]
# tile -> (m, n, k): 1024 tiles of that size, so every CTA walks a stream-K run and then data-parallel tiles
SHAPES = {64: (4096, 1024, 1024), 128: (4096, 4096, 512), 192: (4096, 2048, 768)}  # 192 = 128x64 tiles
TILE_NAMES = {64: "64x64", 128: "128x128", 192: "128x64"}


def _operands(ta, tb, m, n, k, odd_ld, pattern=False):
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    lda, ldb, ldc = ar + odd_ld, br + odd_ld, m  # odd leading dimensions take the cp.async producers
    g = torch.Generator(device="cuda").manual_seed(1234 + 7 * ta + 3 * tb + odd_ld)
    if pattern:
        kt = (torch.arange(k, device="cuda") // 16).to(torch.float64)
        A = torch.zeros(lda * ac, dtype=torch.float64, device="cuda")
        Av = A.view(ac, lda)
        if ta:
            Av[:, :k] = kt          # stored k x m: column i holds k-tile indices down its rows
        else:
            Av[:, :m] = kt[:, None]  # stored m x k: column l is constant
        B = torch.ones(ldb * bc, dtype=torch.float64, device="cuda")
    else:
        A = torch.rand(lda * ac, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
        B = torch.rand(ldb * bc, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    C0 = torch.rand(ldc * n, dtype=torch.float64, device="cuda", generator=g) * 2 - 1
    return A, lda, B, ldb, C0, ldc


def _stress(handle, tile, ta, tb, odd_ld, iters, pattern=False):
    m, n, k = SHAPES[tile]
    A, lda, B, ldb, C0, ldc = _operands(ta, tb, m, n, k, odd_ld, pattern)
    alpha, beta = (1.0, 0.0) if pattern else (1.25, -0.5)
    handle.set_option("dgemm.variant", 3)
    handle.set_option("dgemm.tile", tile)
    try:
        C = C0.clone()
        handle.gemm(True, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
        kernel = handle.last_kernel()
        assert TILE_NAMES[tile] in kernel and ("cpasync" in kernel) == bool(odd_ld), kernel
        if pattern:
            T = k // 16
            ref = torch.full_like(C, 16.0 * T * (T - 1) / 2)
        else:
            ref = C.clone()
        bad = torch.zeros((), dtype=torch.int64, device="cuda")
        for _ in range(iters):
            C.copy_(C0)
            handle.gemm(True, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc)
            bad += (C != ref).any()
        torch.cuda.synchronize()
        assert int(bad.item()) == 0, f"{int(bad.item())} of {iters} runs of {kernel} differ from the first (op {'NT'[ta]}{'NT'[tb]})"
    finally:
        handle.set_option("dgemm.variant", 0)
        handle.set_option("dgemm.tile", 0)


@pytest.mark.parametrize("ta,tb", OPS)
@pytest.mark.parametrize("odd_ld", [0, 1], ids=["tma", "cpasync"])
@pytest.mark.parametrize("tile", [64, 128, 192])
def test_dgemm_repeats_bit_identically(handle, tile, odd_ld, ta, tb):
    _stress(handle, tile, ta, tb, odd_ld, ITERS)


@pytest.mark.parametrize("ta,tb", [(1, 0), (0, 0)])
@pytest.mark.parametrize("tile", [64, 128, 192])
def test_dgemm_k_tile_pattern_is_exact(handle, tile, ta, tb):
    """exact answer known: no run may take a fragment from another k-tile (older or newer fill of its ring stage)"""
    _stress(handle, tile, ta, tb, 0, ITERS, pattern=True)


def test_host_pipeline_repeats_bit_identically(handle):
    """the scenario in which the race was first seen (round 1): the host-buffer pipeline, op TN, ragged shapes, beta != 0"""
    import rtatblas_b200 as rt

    cs = rt.ColumnSplit(handle, 0, 1)
    ta, tb, m, n, k = 1, 0, 1000, 703, 3000
    cs.host_setup(k * m * 8)
    cs.host_connect(None)
    try:
        g = torch.Generator().manual_seed(11)
        hA = torch.rand(k * m, dtype=torch.float64, generator=g).mul_(2).sub_(1).pin_memory()
        hB = torch.rand(n * k, dtype=torch.float64, generator=g).mul_(2).sub_(1).pin_memory()
        ldc = m + 5
        hC0 = torch.rand(ldc * n, dtype=torch.float64, generator=g).pin_memory()
        hC = hC0.clone().pin_memory()
        ref = None
        for rep in range(300):
            hC.copy_(hC0)
            cs.dgemm_host(ta, tb, m, n, k, 1.25, hA, k, hB, k, -0.5, hC, ldc)
            if ref is None:
                ref = hC.clone()
            else:
                assert torch.equal(hC, ref), f"repetition {rep} of the host pipeline differs from the first"
    finally:
        cs.host_teardown()
        cs.close()
