"""-m gpu parity tests for the SYRK and TRSM paths: rtat_b200_{d,s}syrk / _trsm through
the C ABI and the SYRK / TRSM planners through librtat_host.so, against the CPU oracle on the same seeded inputs.
Shapes follow the reference's own tests (tests/executor_test.cpp:132-290: TRSM 435 x 527, SYRK n = 213 k = 74)
plus tile-boundary, ragged, unaligned and degenerate cases.  Tolerances:
  SYRK  BASELINE.json's GEMM bound  |C - C_exact| <= c k eps (|A||A|^T)  on the owned triangle, c = 2 (f64) / 8 (f32);
        everything outside the triangle bit-identical to the input.
  TRSM  componentwise backward-error bound of substitution  |op(A) X - alpha B| <= c na eps (|op(A)||X| + |alpha B|),
        c = 4 (f64) / 16 (f32: 3xTF32 updates), and agreement with the oracle's solution on well-conditioned A."""
import numpy as np
import pytest
import torch

import rtatblas_b200 as rt
from gpu_util import dev, gemm_tolerance, host
from oracle_binding import BOOL_PLANS, view

pytestmark = pytest.mark.gpu

LOWER, UPPER = 0, 1  # RTAT_B200_FILL_*
LEFT, RIGHT = 0, 1   # RTAT_B200_SIDE_*


def _tri_mask(n, lower):
    return np.tril(np.ones((n, n), bool)) if lower else np.triu(np.ones((n, n), bool))


def _syrk_case(oracle, handle, dtype, lower, trans, n, k, alpha, beta, seed, lda_pad=0, ldc_pad=0, offset=0, poison=True):
    ar, ac = (k, n) if trans else (n, k)
    lda, ldc = max(ar, 1) + lda_pad, max(n, 1) + ldc_pad
    A = oracle.fill(lda * ac + offset, seed, dtype)
    C0 = oracle.fill(ldc * n + offset, seed + 1, dtype)
    tri = _tri_mask(n, lower)
    if poison:  # the triangle SYRK does not own must never be read; with beta == 0 neither is the owned one
        view(C0[offset:], n, n, ldc)[~tri] = np.nan
        if beta == 0:
            view(C0[offset:], n, n, ldc)[tri] = np.nan
    dA, dC = dev(A), dev(C0)
    isz = A.itemsize
    handle.syrk(dtype == np.float64, UPPER if not lower else LOWER, int(trans), n, k, alpha, dA.data_ptr() + offset * isz, lda, beta,
                dC.data_ptr() + offset * isz, ldc)
    got = host(dC)
    Cin = None if beta == 0 else np.nan_to_num(C0[offset:])
    ex, ab = oracle.gemm_exact(int(trans), int(not trans), n, n, k, alpha, A[offset:], lda, A[offset:], lda, beta, Cin, ldc)
    tol = gemm_tolerance(dtype, k, ab, ex)
    G = view(got[offset:], n, n, ldc).astype(np.float64)
    err = np.abs(G - ex)
    assert np.all(err[tri] <= tol[tri]), f"max err {np.nanmax(err[tri]):.3e} kernel={handle.last_kernel()}"
    # other triangle, ld padding and leading offset: bit-identical to the input
    keep = np.ones(got.shape, bool)
    view(keep[offset:], n, n, ldc)[tri] = False
    assert np.array_equal(got[keep].view(np.uint8), C0[keep].view(np.uint8)), "wrote outside the owned triangle"
    return handle.last_kernel()


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("lower", [0, 1])
@pytest.mark.parametrize("trans", [0, 1])
def test_syrk_reference_shapes_and_edges(oracle, handle, dtype, lower, trans):
    # executor_test.cpp:230-231 (n = 213, k = 74, alpha -1 beta 1), harness defaults alpha 1 beta 0; tile edges 64 / 128
    _syrk_case(oracle, handle, dtype, lower, trans, 213, 74, -1.0, 1.0, 500)
    _syrk_case(oracle, handle, dtype, lower, trans, 213, 74, 1.0, 0.0, 510)
    for n, k in [(1, 1), (37, 19), (64, 16), (65, 17), (128, 32), (129, 33), (300, 257)]:
        _syrk_case(oracle, handle, dtype, lower, trans, n, k, 0.75, -0.5, 520 + n)
    # operands TMA cannot address: odd leading dimension, base pointer only element-aligned
    _syrk_case(oracle, handle, dtype, lower, trans, 150, 70, 1.25, 0.5, 540, lda_pad=1, ldc_pad=3, offset=1)
    _syrk_case(oracle, handle, dtype, lower, trans, 130, 96, 1.0, 0.0, 550, lda_pad=4, ldc_pad=2)


@pytest.mark.parametrize("tile", [64, 128])
def test_dsyrk_tiles_producers_and_streamk(oracle, handle, tile):
    handle.set_option("dgemm.tile", tile)
    try:
        for lower in (0, 1):
            for trans in (0, 1):
                # k = 1024: 64 k-tiles, few tiles => the stream-K tail with its fix-up slots runs
                name = _syrk_case(oracle, handle, np.float64, lower, trans, 260, 1024, 1.0, 0.0, 600)
                assert name.startswith("dsyrk_ws_dmma") and "tma" in name and "streamk" in name, name
                name = _syrk_case(oracle, handle, np.float64, lower, trans, 260, 1024, -1.0, 1.0, 610, lda_pad=1)
                assert "cpasync8" in name, name
    finally:
        handle.set_option("dgemm.tile", 0)


def test_ssyrk_runs_on_tcgen05(oracle, handle):
    for lower in (0, 1):
        for trans in (0, 1):
            name = _syrk_case(oracle, handle, np.float32, lower, trans, 384, 512, 1.0, 0.0, 650)
            assert name.startswith("ssyrk_tcgen05_3xtf32_2cta"), name   # CTA pairs walk the triangle pair row by pair row
            # even and odd numbers of 128-row tiles, several pair rows, ragged last tile, beta != 0, padded C; and the
            # single-CTA kernel on the same problem
            for n, k, beta in ((1024, 96, 0.0), (1000, 130, -0.5), (641, 64, 1.0)):
                name = _syrk_case(oracle, handle, np.float32, lower, trans, n, k, 1.25, beta, 670 + n, ldc_pad=3)
                assert name.startswith("ssyrk_tcgen05_3xtf32_2cta"), name
            handle.set_option("sgemm.pair", 0)
            try:
                name = _syrk_case(oracle, handle, np.float32, lower, trans, 1000, 130, 1.25, -0.5, 1670, ldc_pad=3)
            finally:
                handle.set_option("sgemm.pair", 1)
            assert name.startswith("ssyrk_tcgen05_3xtf32_128x128"), name
            name = _syrk_case(oracle, handle, np.float32, lower, trans, 200, 96, 1.0, 0.5, 660, lda_pad=1)
            assert name.startswith("ssyrk_tcgen05_3xtf32") and name.endswith("_padA"), name  # odd ld: aligned copy of A first
            handle.set_option("sgemm.pad_a", 0)
            try:
                name = _syrk_case(oracle, handle, np.float32, lower, trans, 200, 96, 1.0, 0.5, 660, lda_pad=1)
            finally:
                handle.set_option("sgemm.pad_a", 1)
            assert name.startswith("ssyrk_simt"), name


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_syrk_degenerate_and_argument_checks(oracle, handle, dtype):
    dbl = dtype == np.float64
    n = 70
    for lower in (0, 1):
        tri = _tri_mask(n, lower)
        for alpha, k, beta in [(0.0, 9, 0.5), (1.0, 0, -2.0), (0.0, 9, 0.0), (0.0, 9, 1.0)]:
            C0 = oracle.fill((n + 2) * n, 700, dtype)
            if beta == 0:
                view(C0, n, n, n + 2)[tri] = np.nan  # beta == 0: cleared without being read
            dC = dev(C0)
            A = dev(np.full(max(n * max(k, 1), 1), np.nan, dtype))  # alpha == 0 or k == 0: A is not referenced
            handle.syrk(dbl, UPPER if not lower else LOWER, 0, n, k, alpha, A, n, beta, dC, n + 2)
            got = host(dC)
            exp = C0.copy()
            V = view(exp, n, n, n + 2)
            V[tri] = 0 if beta == 0 else (V[tri] * dtype(beta)).astype(dtype)
            assert np.array_equal(got.view(np.uint8), exp.view(np.uint8))
    x = dev(np.zeros(16, dtype))
    assert handle.syrk(dbl, LOWER, 0, 0, 5, 1.0, x, 1, 0.0, x, 1) == 0  # n == 0: nothing to do
    INVALID = 7
    assert handle.syrk(dbl, 2, 0, 4, 4, 1.0, x, 4, 0.0, x, 4, check=False) == INVALID   # bad uplo
    assert handle.syrk(dbl, LOWER, 2, 4, 4, 1.0, x, 4, 0.0, x, 4, check=False) == INVALID  # bad trans
    assert handle.syrk(dbl, LOWER, 0, 4, 2, 1.0, x, 3, 0.0, x, 4, check=False) == INVALID  # lda < n
    assert handle.syrk(dbl, LOWER, 1, 4, 2, 1.0, x, 1, 0.0, x, 4, check=False) == INVALID  # lda < k (trans)
    assert handle.syrk(dbl, LOWER, 0, 4, 2, 1.0, x, 4, 0.0, x, 3, check=False) == INVALID  # ldc < n
    assert handle.syrk(dbl, LOWER, 0, -1, 2, 1.0, x, 4, 0.0, x, 4, check=False) == INVALID


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_syrk_plans_through_planner_match_oracle(oracle, handle, dtype):
    """Every SYRK_Options plan via Planning_System<SYRK_Executor> vs the oracle's restatement of form_operation:
    owned triangle within the GEMM bound, the other triangle bit-exact (untouched, or beta-scaled by the
    transpose_C plans — the reference's observable quirk)."""
    planner = rt.Planner("syrk", "double" if dtype == np.float64 else "float")
    assert planner.plans() == BOOL_PLANS
    n, k, alpha, beta = 213, 74, -1.0, 0.5
    for lower in (0, 1):
        for trans in (0, 1):
            ar, ac = (k, n) if trans else (n, k)
            A = oracle.fill(ar * ac, 800, dtype)
            C0 = oracle.fill(n * n, 801, dtype)
            tri = _tri_mask(n, lower)
            ex, ab = oracle.gemm_exact(trans, 1 - trans, n, n, k, alpha, A, ar, A, ar, beta, C0, n)
            tol = gemm_tolerance(dtype, k, ab, ex) + np.finfo(dtype).eps * np.abs(ex)  # one more rounding in the accumulate
            for plan in BOOL_PLANS:
                want = C0.copy()
                rc, _ = oracle.syrk_plan(plan, lower, trans, n, k, alpha, A, ar, beta, want, n)
                assert rc == 0
                need = planner.syrk_workspace(UPPER if not lower else LOWER, trans, n, k, plan)
                assert need == oracle.syrk_plan_workspace(plan, trans, n, k, A.itemsize)
                ws = torch.full((need + 16,), 0xFF, dtype=torch.uint8, device="cuda")  # NaN-patterned scratch
                dA, dC = dev(A), dev(C0)
                used = planner.syrk(handle, UPPER if not lower else LOWER, trans, n, k, alpha, dA, ar, beta, dC, n, ws, need, plan=plan)
                assert used == plan
                got = view(host(dC), n, n, n)
                W = view(want, n, n, n)
                assert np.all(np.abs(got.astype(np.float64) - ex)[tri] <= tol[tri]), (plan, lower, trans)
                assert np.array_equal(got[~tri].view(np.uint8), W[~tri].view(np.uint8)), (plan, lower, trans)
                assert bool((ws[need:] == 0xFF).all()), "wrote past the plan's workspace"
    # autotune on a key nothing above has sampled: four exploration calls, then a converged plan that is one of the four
    n2, k2 = 150, 40
    dA, dC = dev(oracle.fill(n2 * k2, 810, dtype)), dev(oracle.fill(n2 * n2, 811, dtype))
    ws = torch.empty(planner.syrk_workspace(LOWER, 0, n2, k2, "TT") + 16, dtype=torch.uint8, device="cuda")
    seen = [planner.syrk(handle, LOWER, 0, n2, k2, 1.0, dA, n2, 0.0, dC, n2, ws, ws.numel()) for _ in range(6)]
    assert seen[:4] == BOOL_PLANS and seen[4] == seen[5] and seen[4] in BOOL_PLANS
    stats = planner.statistics()
    assert {"uplo": "Lower", "trans": "N", "n": n2, "k": k2} in [p["key"] for p in stats]
    planner.close()


def test_dsyrk_full_size_triangle_matches_gemm(handle):
    """n = k = 4096 (C1/C2-class sizes): the SYRK triangle against cuBLAS' product (torch, test-only use), the other
    triangle untouched; also checks the schedule really skipped: about half the GEMM's time is not asserted,
    but the kernel name and tile count are."""
    n = k = 4096
    A = torch.empty(n * k, dtype=torch.float64, device="cuda")
    handle.fill_uniform(A, 9100)
    for lower, trans in [(1, 0), (0, 1)]:
        Cm = torch.full((n * n,), float("nan"), dtype=torch.float64, device="cuda")
        handle.syrk(True, LOWER if lower else UPPER, trans, n, k, 1.0, A, n if not trans else k, 0.0, Cm, n)
        assert handle.last_kernel().startswith("dsyrk_ws_dmma_128x128x16_tma"), handle.last_kernel()
        Am = A.view(k, n).t() if not trans else A.view(n, k).t()  # stored (n x k) or (k x n), column-major
        opA = Am if not trans else Am.t()
        ref = opA @ opA.t()
        bound = opA.abs() @ opA.abs().t()
        got = Cm.view(n, n).t()
        tri = torch.tril(torch.ones(n, n, dtype=torch.bool, device="cuda")) if lower else torch.triu(torch.ones(n, n, dtype=torch.bool, device="cuda"))
        eps = np.finfo(np.float64).eps
        tol = 2 * 2 * k * eps * bound + eps * ref.abs()
        assert bool(torch.all((got - ref).abs()[tri] <= tol[tri]))
        assert bool(torch.isnan(got[~tri]).all()), "SYRK touched the other triangle"


# ---------------------------------------------------------------------------------------------------------------
# TRSM
# ---------------------------------------------------------------------------------------------------------------
def _triangular(oracle, dtype, na, lda, seed, unit, offset=0):
    """Uniform fill; diagonal pushed away from zero (non-unit) or off-diagonals shrunk (unit), like
    executor_test.cpp:146-176.  The unused triangle keeps its random contents."""
    A = oracle.fill(lda * na + offset, seed, dtype)
    V = view(A[offset:], na, na, lda)
    d = np.arange(na)
    diag = V[d, d].copy()
    V[:] = V / dtype(4)
    V[d, d] = diag + np.where(diag < 0, -2, 2).astype(dtype)
    return A


def _trsm_check(dtype, left, lower, trans, unit, m, n, alpha, A, lda, B0, got, ldb, oracle_x=None, what=""):
    na = m if left else n
    T = (np.tril(view(A, na, na, lda)) if lower else np.triu(view(A, na, na, lda))).astype(np.float64)
    if unit:
        np.fill_diagonal(T, 1.0)
    op = T.T if trans else T
    X = view(got, m, n, ldb).astype(np.float64)
    rhs = alpha * view(B0, m, n, ldb).astype(np.float64)
    res = (op @ X - rhs) if left else (X @ op - rhs)
    bound = (np.abs(op) @ np.abs(X) + np.abs(rhs)) if left else (np.abs(X) @ np.abs(op) + np.abs(rhs))
    c = 4 if dtype == np.float64 else 16
    tol = c * na * np.finfo(dtype).eps * bound
    assert np.all(np.abs(res) <= tol), f"{what} residual {np.abs(res).max():.3e} vs {tol.flat[np.abs(res).argmax()]:.3e}"
    if oracle_x is not None:
        Xo = view(oracle_x, m, n, ldb).astype(np.float64)
        rel = np.abs(X - Xo).max() / max(np.abs(Xo).max(), 1e-300)
        assert rel < (1e-10 if dtype == np.float64 else 2e-4), f"{what} differs from the oracle by {rel:.3e}"


def _trsm_case(oracle, handle, dtype, left, lower, trans, unit, m, n, alpha, seed, lda_pad=0, ldb_pad=0, offset=0, compare=True):
    na = m if left else n
    lda, ldb = na + lda_pad, m + ldb_pad
    A = _triangular(oracle, dtype, na, lda, seed, unit, offset)
    B0 = oracle.fill(ldb * n + offset, seed + 1, dtype)
    want = None
    if compare:
        want = B0.copy()
        oracle.trsm(int(left), int(lower), int(trans), int(unit), m, n, alpha, A[offset:], lda, want[offset:], ldb)
    # poison what TRSM must not read: the other triangle, and the diagonal of a unit-diagonal A
    Ap = A.copy()
    V = view(Ap[offset:], na, na, lda)
    other = ~_tri_mask(na, lower)
    V[other] = np.nan
    if unit:
        V[np.arange(na), np.arange(na)] = np.nan
    dA, dB = dev(Ap), dev(B0)
    isz = A.itemsize
    handle.trsm(dtype == np.float64, RIGHT if not left else LEFT, UPPER if not lower else LOWER, int(trans), int(unit), m, n, alpha,
                dA.data_ptr() + offset * isz, lda, dB.data_ptr() + offset * isz, ldb)
    got = host(dB)
    _trsm_check(dtype, left, lower, trans, unit, m, n, alpha, A[offset:], lda, B0[offset:], got[offset:], ldb,
                None if want is None else want[offset:], what=f"left={left} lower={lower} trans={trans} unit={unit} m={m} n={n}")
    keep = np.ones(got.shape, bool)
    view(keep[offset:], m, n, ldb)[:] = False
    assert np.array_equal(got[keep].view(np.uint8), B0[keep].view(np.uint8)), "wrote outside B's m x n window"


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("left", [0, 1])
@pytest.mark.parametrize("lower", [0, 1])
def test_trsm_all_variants_small_and_ragged(oracle, handle, dtype, left, lower):
    for trans in (0, 1):
        for unit in (0, 1):
            for (m, n) in [(1, 1), (21, 17), (32, 40), (33, 31), (64, 64), (100, 7), (7, 100), (129, 67)]:
                _trsm_case(oracle, handle, dtype, left, lower, trans, unit, m, n, 1.0, 1000 + m)
            _trsm_case(oracle, handle, dtype, left, lower, trans, unit, 75, 90, -1.5, 1100, lda_pad=1, ldb_pad=3, offset=1)


@pytest.mark.parametrize("left", [0, 1])
@pytest.mark.parametrize("lower", [0, 1])
@pytest.mark.parametrize("trans", [0, 1])
def test_dtrsm_reference_test_shape(oracle, handle, left, lower, trans):
    # executor_test.cpp:135-136: m = 435, n = 527
    for unit in (0, 1):
        _trsm_case(oracle, handle, np.float64, left, lower, trans, unit, 435, 527, 1.0, 1200)
    assert handle.last_kernel().startswith("dtrsm_recursive")


@pytest.mark.parametrize("base", [32, 64, 256])
def test_dtrsm_recursion_cutoffs(oracle, handle, base):
    handle.set_option("trsm.base", base)
    try:
        for left, lower, trans, unit in [(1, 1, 0, 0), (1, 0, 1, 1), (0, 1, 1, 0), (0, 0, 0, 1)]:
            _trsm_case(oracle, handle, np.float64, left, lower, trans, unit, 300, 277, -0.5, 1250)
    finally:
        handle.set_option("trsm.base", 0)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_syrk_trsm_extreme_aspect_ratios(oracle, handle, dtype):
    """One tile with a very long k (the whole stream-K grid feeds it), a rank-1/2 update of a large triangle, a single
    right-hand side against a deep recursion, and a 1 x 1 triangle against thousands of right-hand sides."""
    _syrk_case(oracle, handle, dtype, 1, 0, 64, 60000, 1.0, 0.0, 2000)
    _syrk_case(oracle, handle, dtype, 0, 1, 96, 40001, -1.0, 1.0, 2001)
    _syrk_case(oracle, handle, dtype, 1, 1, 1500, 2, 0.5, 2.0, 2002)
    _syrk_case(oracle, handle, dtype, 0, 0, 1300, 1, 1.0, 0.0, 2003)
    _trsm_case(oracle, handle, dtype, 1, 1, 0, 0, 1100, 1, 1.0, 2010)
    _trsm_case(oracle, handle, dtype, 0, 0, 1, 1, 1, 1100, -2.0, 2011)
    _trsm_case(oracle, handle, dtype, 1, 0, 1, 0, 1, 5000, 1.0, 2012)
    _trsm_case(oracle, handle, dtype, 0, 1, 0, 1, 5000, 1, 0.5, 2013)


def test_strsm_mid_size(oracle, handle):
    for left, lower, trans in [(1, 1, 0), (0, 0, 1), (1, 0, 0), (0, 1, 1)]:
        _trsm_case(oracle, handle, np.float32, left, lower, trans, 0, 520, 384, 1.0, 1300)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_trsm_degenerate_and_argument_checks(oracle, handle, dtype):
    dbl = dtype == np.float64
    B0 = oracle.fill(12 * 9, 1400, dtype)
    dB = dev(B0)
    A = dev(np.full(100, np.nan, dtype))
    handle.trsm(dbl, LEFT, LOWER, 0, 0, 10, 9, 0.0, A, 10, dB, 12)  # alpha == 0: B = 0, A not referenced
    got = view(host(dB), 12, 9, 12)
    assert np.all(got[:10] == 0) and np.array_equal(got[10:], view(B0, 12, 9, 12)[10:])
    x = dev(np.zeros(16, dtype))
    assert handle.trsm(dbl, LEFT, LOWER, 0, 0, 0, 4, 1.0, x, 1, x, 1) == 0
    assert handle.trsm(dbl, LEFT, LOWER, 0, 0, 4, 0, 1.0, x, 4, x, 4) == 0
    INVALID = 7
    assert handle.trsm(dbl, 2, LOWER, 0, 0, 4, 4, 1.0, x, 4, x, 4, check=False) == INVALID
    assert handle.trsm(dbl, LEFT, 2, 0, 0, 4, 4, 1.0, x, 4, x, 4, check=False) == INVALID
    assert handle.trsm(dbl, LEFT, LOWER, 2, 0, 4, 4, 1.0, x, 4, x, 4, check=False) == INVALID
    assert handle.trsm(dbl, LEFT, LOWER, 0, 2, 4, 4, 1.0, x, 4, x, 4, check=False) == INVALID
    assert handle.trsm(dbl, LEFT, LOWER, 0, 0, 4, 2, 1.0, x, 3, x, 4, check=False) == INVALID   # lda < m (left)
    assert handle.trsm(dbl, RIGHT, LOWER, 0, 0, 2, 4, 1.0, x, 3, x, 2, check=False) == INVALID  # lda < n (right)
    assert handle.trsm(dbl, LEFT, LOWER, 0, 0, 4, 2, 1.0, x, 4, x, 3, check=False) == INVALID   # ldb < m


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_trsm_plans_through_planner_match_oracle(oracle, handle, dtype):
    """Every TRSM_Options plan via Planning_System<TRSM_Executor> vs the oracle's restatement of form_operation."""
    planner = rt.Planner("trsm", "double" if dtype == np.float64 else "float")
    assert planner.plans() == BOOL_PLANS
    m, n, alpha = 131, 97, 1.0
    for left in (0, 1):
        for lower in (0, 1):
            for trans in (0, 1):
                for unit in (0, 1):
                    na = m if left else n
                    A = _triangular(oracle, dtype, na, na, 1500, unit)
                    B0 = oracle.fill(m * n, 1501, dtype)
                    for plan in BOOL_PLANS:
                        want = B0.copy()
                        rc, _ = oracle.trsm_plan(plan, left, lower, trans, unit, m, n, alpha, A, na, want, m)
                        assert rc == 0
                        args = (RIGHT if not left else LEFT, UPPER if not lower else LOWER, trans, unit, m, n)
                        need = planner.trsm_workspace(*args, plan)
                        assert need == oracle.trsm_plan_workspace(plan, left, m, n, A.itemsize)
                        ws = torch.full((need + 16,), 0xFF, dtype=torch.uint8, device="cuda")
                        dA, dB = dev(A), dev(B0)
                        assert planner.trsm(handle, *args, alpha, dA, na, dB, m, ws, need, plan=plan) == plan
                        _trsm_check(dtype, left, lower, trans, unit, m, n, alpha, A, na, B0, host(dB), m, want,
                                    what=f"plan={plan} left={left} lower={lower} trans={trans} unit={unit}")
                        assert bool((ws[need:] == 0xFF).all()), "wrote past the plan's workspace"
    stats_key = {"side": "Left", "uplo": "Lower", "trans": "N", "diag": "Non-Unit", "m": 64, "n": 48}
    dA, dB = dev(_triangular(oracle, dtype, 64, 64, 1510, 0)), dev(oracle.fill(64 * 48, 1511, dtype))
    ws = torch.empty(planner.trsm_workspace(LEFT, LOWER, 0, 0, 64, 48, "TT") + 16, dtype=torch.uint8, device="cuda")
    seen = [planner.trsm(handle, LEFT, LOWER, 0, 0, 64, 48, 1.0, dA, 64, dB, 64, ws, ws.numel()) for _ in range(6)]
    assert seen[:4] == BOOL_PLANS and seen[4] == seen[5]
    assert stats_key in [p["key"] for p in planner.statistics()]
    planner.close()


def test_dtrsm_full_size_residual(handle):
    """m = n = 4096, left / lower / N and right / upper / T: the recursion reaches the 128 x 128 DMMA kernel with k up to
    2048; checked through the size-independent property  op(A) X == alpha B  (residual bound above) on device."""
    n = 4096
    A = torch.empty(n * n, dtype=torch.float64, device="cuda")
    B0 = torch.empty(n * n, dtype=torch.float64, device="cuda")
    handle.fill_uniform(A, 9200)
    handle.fill_uniform(B0, 9201)
    Am = A.view(n, n).t()  # column-major view
    Am.mul_(1.0 / 64).diagonal().add_(2.0)
    for left, lower, trans in [(1, 1, 0), (0, 0, 1)]:
        B = B0.clone()
        launches = handle.launch_count()
        handle.trsm(True, LEFT if left else RIGHT, LOWER if lower else UPPER, trans, 0, n, n, 1.0, A, n, B, n)
        assert handle.launch_count() - launches >= 2 * (n // 128) - 1  # 32 block solves + 31 GEMM updates
        T = torch.tril(Am) if lower else torch.triu(Am)
        op = T.t() if trans else T
        X, R = B.view(n, n).t(), B0.view(n, n).t()
        res = (op @ X - R) if left else (X @ op - R)
        bound = (op.abs() @ X.abs() + R.abs()) if left else (X.abs() @ op.abs() + R.abs())
        tol = 4 * n * np.finfo(np.float64).eps * bound
        assert bool(torch.all(res.abs() <= tol)), f"residual {res.abs().max().item():.3e}"
