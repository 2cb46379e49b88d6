"""-m gpu parity tests: the sm_100a kernels, called through the C ABI (include/rtat_b200.h), against
the CPU oracle on the same seeded inputs.  geam is bit-exact; gemm is held to BASELINE.json's bound.
Shapes follow the reference's own tests (tests/executor_test.cpp, matrixop_test.cpp,
planning_test.cpp) plus tile-boundary, ragged, unaligned and degenerate cases."""
import numpy as np
import pytest
import torch

import rtatblas_b200 as rt
import oracle_binding
from gpu_util import dev, gemm_tolerance, host, sampled_exact_ratio
from oracle_binding import view

pytestmark = pytest.mark.gpu

_ORACLE = []


def ORACLE():
    """the CPU oracle for helpers that are not handed the fixture (loaded once)"""
    if not _ORACLE:
        _ORACLE.append(oracle_binding.load())
    return _ORACLE[0]


def _gemm_case(oracle, handle, dtype, ta, tb, m, n, k, alpha, beta, seed, lda_pad=0, ldb_pad=0, ldc_pad=0, offset=0,
               nan_c=False):
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    lda, ldb, ldc = max(ar, 1) + lda_pad, max(br, 1) + ldb_pad, max(m, 1) + ldc_pad
    A = oracle.fill(lda * ac + offset, seed, dtype)
    B = oracle.fill(ldb * bc + offset, seed + 1, dtype)
    C0 = oracle.fill(ldc * n + offset, seed + 2, dtype)
    if nan_c:
        C0[:] = np.nan
    dA, dB, dC = dev(A), dev(B), dev(C0)
    isz = A.itemsize
    handle.gemm(dtype == np.float64, int(ta), int(tb), m, n, k, alpha, dA.data_ptr() + offset * isz, lda,
                dB.data_ptr() + offset * isz, ldb, beta, dC.data_ptr() + offset * isz, ldc)
    got = host(dC)
    ex, ab = oracle.gemm_exact(int(ta), int(tb), m, n, k, alpha, A[offset:], lda, B[offset:], ldb, beta, C0[offset:], ldc)
    tol = gemm_tolerance(dtype, k, ab, ex)
    err = np.abs(view(got[offset:], m, n, ldc).astype(np.float64) - ex)
    assert np.all(err <= tol), f"max err {err.max():.3e} vs tol {tol.flat[err.argmax()]:.3e} kernel={handle.last_kernel()}"
    # everything outside the m x n window (ld padding, leading offset) is untouched
    mask = np.ones(got.shape, bool)
    vm = view(mask[offset:], m, n, ldc)
    vm[:] = False
    assert np.array_equal(got[mask].view(np.uint8), C0[mask].view(np.uint8)), "wrote outside C's m x n window"
    return err.max()


OPS = [(0, 0), (0, 1), (1, 0), (1, 1)]


@pytest.mark.parametrize("ta,tb", OPS)
@pytest.mark.parametrize("variant", [1, 2, 3, 4])
def test_dgemm_reference_test_shapes(oracle, handle, ta, tb, variant):
    handle.set_option("dgemm.variant", variant)
    try:
        # executor_test.cpp:15-17, planning_test.cpp:89-91/:121-123/:151-153, matrixop_test.cpp:34-36
        for (m, n, k) in [(70, 45, 62), (23, 16, 35), (423, 125, 318), (69, 123, 42), (26, 27, 34)]:
            _gemm_case(oracle, handle, np.float64, ta, tb, m, n, k, 1.0, 0.0, 100 + m)
            _gemm_case(oracle, handle, np.float64, ta, tb, m, n, k, -0.75, 0.5, 200 + m, lda_pad=3, ldb_pad=1, ldc_pad=2)
    finally:
        handle.set_option("dgemm.variant", 0)


@pytest.mark.parametrize("ta,tb", OPS)
@pytest.mark.parametrize("variant,tile", [(0, 0), (3, 128), (3, 64), (3, 192)])
def test_dgemm_tile_boundaries_and_alignment(oracle, handle, ta, tb, variant, tile):
    handle.set_option("dgemm.variant", variant)
    handle.set_option("dgemm.tile", tile)
    try:
        _dgemm_tile_boundaries_and_alignment(oracle, handle, ta, tb)
    finally:
        handle.set_option("dgemm.variant", 0)
        handle.set_option("dgemm.tile", 0)


def _dgemm_tile_boundaries_and_alignment(oracle, handle, ta, tb):
    for (m, n, k) in [(128, 128, 16), (129, 127, 17), (255, 257, 33), (1, 1, 1), (1, 300, 5), (300, 1, 7), (64, 64, 1),
                      (130, 66, 129), (256, 384, 512)]:
        _gemm_case(oracle, handle, np.float64, ta, tb, m, n, k, 1.0, 0.0, 300 + m + n)
    # odd leading dimensions (config 3 style) and a base pointer that is only 8 B aligned
    _gemm_case(oracle, handle, np.float64, ta, tb, 257, 131, 65, 1.0, 0.0, 41, lda_pad=1, ldb_pad=1, ldc_pad=1)
    _gemm_case(oracle, handle, np.float64, ta, tb, 200, 136, 72, 2.0, -1.0, 43, offset=1)
    handle.set_option("dgemm.force_unaligned", 1)
    try:
        _gemm_case(oracle, handle, np.float64, ta, tb, 256, 128, 64, 1.0, 0.0, 47)
    finally:
        handle.set_option("dgemm.force_unaligned", 0)


def test_dgemm_beta_zero_never_reads_c(oracle, handle):
    # the reference hands uninitialised scratch as C with beta = 0 (matrixop.h:433-436)
    for ta, tb in OPS:
        _gemm_case(oracle, handle, np.float64, ta, tb, 150, 90, 40, 1.0, 0.0, 51, nan_c=True)


def test_gemm_degenerate_cases(oracle, handle):
    for dtype in (np.float64, np.float32):
        # k == 0 and alpha == 0: C = beta*C without touching A/B
        _gemm_case(oracle, handle, dtype, 0, 0, 37, 19, 0, 1.0, 0.5, 61)
        _gemm_case(oracle, handle, dtype, 0, 0, 37, 19, 0, 1.0, 0.0, 62, nan_c=True)
        _gemm_case(oracle, handle, dtype, 1, 0, 37, 19, 11, 0.0, -2.0, 63)
        _gemm_case(oracle, handle, dtype, 0, 1, 37, 19, 11, 0.0, 1.0, 64)
    one = 1.0
    # m == 0 / n == 0 are no-ops; bad arguments give INVALID_VALUE (7), not a crash
    assert handle.gemm(True, 0, 0, 0, 5, 5, one, None, 1, None, 5, one, None, 1, check=False) == 0
    assert handle.gemm(True, 0, 0, 5, 0, 5, one, None, 5, None, 5, one, None, 5, check=False) == 0
    x = torch.zeros(64, dtype=torch.float64, device="cuda")
    assert handle.gemm(True, 0, 0, 8, 8, 8, one, x, 4, x, 8, one, x, 8, check=False) == 7   # lda < m
    assert handle.gemm(True, 2, 0, 8, 8, 8, one, x, 8, x, 8, one, x, 8, check=False) == 7   # bad op
    assert handle.gemm(True, 0, 0, -1, 8, 8, one, x, 8, x, 8, one, x, 8, check=False) == 7
    assert handle.gemm(True, 0, 0, 8, 8, 8, one, None, 8, x, 8, one, x, 8, check=False) == 7  # null A


@pytest.mark.parametrize("ta,tb", OPS)
def test_sgemm_reference_test_shapes(oracle, handle, ta, tb):
    # executor_test.cpp:90-130 (the reference demands 1e-4 absolute at k = 62)
    for (m, n, k) in [(70, 45, 62), (23, 16, 35), (129, 127, 65), (1, 1, 1), (260, 130, 40)]:
        err = _gemm_case(oracle, handle, np.float32, ta, tb, m, n, k, 1.0, 0.0, 400 + m)
        if k <= 62:
            assert err <= 1e-4
        _gemm_case(oracle, handle, np.float32, ta, tb, m, n, k, 1.5, -0.5, 500 + m, lda_pad=1, ldb_pad=2, ldc_pad=3)


@pytest.mark.parametrize("variant,tile", [(0, 0), (3, 128), (4, 128), (3, 64), (4, 64), (3, 192), (4, 192)])
def test_dgemm_mid_size_against_oracle(oracle, handle, variant, tile):
    # large enough for many CTAs and k-tiles, small enough for the long-double oracle (~1 s)
    handle.set_option("dgemm.variant", variant)
    handle.set_option("dgemm.tile", tile)
    try:
        _gemm_case(oracle, handle, np.float64, 1, 0, 512, 384, 640, 1.0, 0.0, 71)
        _gemm_case(oracle, handle, np.float64, 0, 1, 513, 383, 641, 1.0, 1.0, 73, lda_pad=1)
        _gemm_case(oracle, handle, np.float64, 0, 0, 1100, 2300, 200, 1.0, 0.0, 75)   # > 148 tiles: persistent loop
        _gemm_case(oracle, handle, np.float64, 1, 1, 2300, 1100, 200, -1.0, 0.5, 77, ldb_pad=2)
    finally:
        handle.set_option("dgemm.variant", 0)
        handle.set_option("dgemm.tile", 0)


# ---------------------------------------------------------------------------------------------
# geam: bit-exact against the oracle (which is pinned on cuBLAS geam outputs, tests/golden)
# ---------------------------------------------------------------------------------------------
def _geam_case(oracle, handle, dtype, ta, tb, m, n, alpha, beta, seed, lda_pad=0, ldb_pad=0, ldc_pad=0, inplace=False,
               offset=0, nan_b=False):
    ar, ac = (n, m) if ta else (m, n)
    br, bc = (n, m) if tb else (m, n)
    lda, ldb, ldc = ar + lda_pad, br + ldb_pad, m + ldc_pad
    A = oracle.fill(lda * ac + offset, seed, dtype)
    A[offset] = -0.0
    B = oracle.fill(ldb * bc + offset, seed + 1, dtype)
    if nan_b:
        B[:] = np.nan
    isz = A.itemsize
    dA = dev(A)
    if inplace:
        assert not tb and ldb == ldc
        C0 = B
        dC = dev(C0)
        dB = dC
    else:
        C0 = oracle.fill(ldc * n + offset, seed + 2, dtype)
        dB, dC = dev(B), dev(C0)
    handle.geam(dtype == np.float64, int(ta), int(tb), m, n, alpha, dA.data_ptr() + offset * isz, lda, beta,
                dB.data_ptr() + offset * isz, ldb, dC.data_ptr() + offset * isz, ldc)
    got = host(dC)
    exp = C0.copy()
    Bh = exp if inplace else B
    oracle.geam(int(ta), int(tb), m, n, dtype(alpha), A[offset:], lda, dtype(beta), Bh[offset:], ldb, exp[offset:], ldc)
    assert np.array_equal(got.view(np.uint8), exp.view(np.uint8)), f"geam mismatch kernel={handle.last_kernel()}"


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_geam_move_shapes(oracle, handle, dtype):
    # MatrixMove = geam(alpha, op(A), beta = 0, C aliases B) (matrixop.h:329-343); MoveTest 10x12 pad 32
    _geam_case(oracle, handle, dtype, 0, 0, 10, 12, 1.0, 0.0, 1, ldb_pad=22, ldc_pad=22, inplace=True, nan_b=True)
    _geam_case(oracle, handle, dtype, 1, 0, 25, 14, -2.0, 0.0, 2, ldb_pad=7, ldc_pad=7, inplace=True, nan_b=True)
    _geam_case(oracle, handle, dtype, 1, 0, 23, 42, 0.5, 0.0, 3, ldb_pad=9, ldc_pad=9, inplace=True, nan_b=True)
    for (m, n) in [(1, 1), (64, 64), (65, 63), (1, 777), (777, 1), (300, 517), (1024, 96), (130, 2050)]:
        for ta in (0, 1):
            _geam_case(oracle, handle, dtype, ta, 0, m, n, 1.0, 0.0, 10 + m, inplace=True, nan_b=True)
            _geam_case(oracle, handle, dtype, ta, 0, m, n, 1.0, 0.0, 11 + m, lda_pad=1, ldb_pad=3, ldc_pad=3, inplace=True, nan_b=True)
            _geam_case(oracle, handle, dtype, ta, 0, m, n, 1.0, 0.0, 12 + m, offset=1, inplace=True, nan_b=True)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_geam_accumulate_and_general(oracle, handle, dtype):
    # MatrixAccumulate = in-place geam(1, op(S), beta, C) (matrixop.h:291-304)
    for beta in (0.0, 1.0, 0.5, -1.25):
        for ta in (0, 1):
            _geam_case(oracle, handle, dtype, ta, 0, 29, 18, 1.0, beta, 20, lda_pad=2, ldb_pad=1, ldc_pad=1, inplace=True)
            _geam_case(oracle, handle, dtype, ta, 0, 333, 129, 1.0, beta, 21, inplace=True)
            _geam_case(oracle, handle, dtype, ta, 0, 520, 388, 1.0, beta, 22, inplace=True)          # TMA transpose(+accumulate) path
            _geam_case(oracle, handle, dtype, ta, 0, 516, 260, -0.5, beta, 23, lda_pad=4, ldb_pad=8, ldc_pad=8, inplace=True)
    # all four op pairs out of place, general alpha/beta, alpha == 0
    for ta, tb in OPS:
        for alpha, beta in [(0.3, -1.7), (-2.0, 1.0), (0.0, 0.6), (1.0, 0.0)]:
            _geam_case(oracle, handle, dtype, ta, tb, 13, 7, alpha, beta, 30, lda_pad=1, ldb_pad=2, ldc_pad=3)
            _geam_case(oracle, handle, dtype, ta, tb, 150, 210, alpha, beta, 31)


def test_geam_transpose_paths_agree_bitwise(oracle, handle):
    """The three transposing kernels — register-blocked (the default), TMA pipeline, scalar tile — on a shape with ragged edge
    tiles, either CTA order and tile height of the default kernel: bit-identical, and equal to the transpose."""
    for dtype in (np.float64, np.float32):
        m, n = 1000, 1100
        A = dev(oracle.fill(n * m, 5, dtype))
        outs = []
        try:
            for sel, order, upc in ((0, 0, 0), (2, 1, 0), (2, 0, 2), (1, 0, 0), (3, 0, 0)):
                handle.set_option("geam.transpose", sel)
                handle.set_option("geam.order", order)
                handle.set_option("geam.upc", upc)
                Cm = torch.full((m * n,), float("nan"), dtype=torch.float64 if dtype == np.float64 else torch.float32, device="cuda")
                handle.geam(dtype == np.float64, 1, 0, m, n, 1.0, A, n, 0.0, Cm, m, Cm, m)
                outs.append((host(Cm), handle.last_kernel()))
        finally:
            for key in ("geam.transpose", "geam.order", "geam.upc"):
                handle.set_option(key, 0)
        assert [o[1].rsplit("_v16", 2)[0] for o in outs] == ["geam_transpose_reg"] * 3 + ["geam_transpose_tma", "geam_tile_tn"], [o[1] for o in outs]
        for o in outs[1:]:
            assert np.array_equal(outs[0][0].view(np.uint8), o[0].view(np.uint8)), o[1]
        assert np.array_equal(outs[0][0].reshape(n, m), host(A).reshape(m, n).T)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_geam_transpose_of_unaligned_source_into_padded_target(oracle, handle, dtype):
    """Config 3's transposing pad plans: A has an odd leading dimension (TMA cannot address it), the padded target can be
    TMA-stored.  cp.async producers fill the TMA kernel's ring; bit-exact against the oracle for move (beta = 0, NaN in the
    target) and accumulate, ragged edge tiles, an element-aligned base pointer, and against the all-TMA path."""
    ld_unit = 16 // np.dtype(dtype).itemsize
    try:
        for sel, cp_name, tma_name in ((0, "geam_transpose_reg_s_v16", "geam_transpose_reg_v16"), (1, "geam_transpose_cpasync_tma", "geam_transpose_tma")):
            handle.set_option("geam.transpose", sel)
            _unaligned_source_cases(oracle, handle, dtype, ld_unit, cp_name, tma_name)
    finally:
        handle.set_option("geam.transpose", 0)


def _unaligned_source_cases(oracle, handle, dtype, ld_unit, cp_name, tma_name):
    for (m, n, lda_pad, beta, offset) in [(517, 300, 1, 0.0, 0), (300, 517, 0, 0.0, 0), (1024, 257, 0, -0.75, 0), (640, 512, 0, 0.0, 1),
                                          (2049, 1301, 0, 1.0, 0)]:
        ldc_pad = (-m) % 32 if offset == 0 else 0
        if offset:  # base moved by one element: the target must stay 16 B aligned, so only A is offset here
            ar, ac = n, m
            lda, ldc = ar + lda_pad, m
            A = oracle.fill(lda * ac + 1, 90 + m, dtype)
            C0 = np.full(ldc * n, np.nan, dtype=dtype)
            dA, dC = dev(A), dev(C0)
            handle.geam(dtype == np.float64, 1, 0, m, n, 1.0, dA.data_ptr() + A.itemsize, lda, 0.0, dC, ldc, dC, ldc)
            assert handle.last_kernel().startswith(cp_name), handle.last_kernel()
            exp = C0.copy()
            oracle.geam(1, 0, m, n, dtype(1.0), A[1:], lda, dtype(0.0), exp, ldc, exp, ldc)
            assert np.array_equal(host(dC).view(np.uint8), exp.view(np.uint8))
            continue
        assert (n + lda_pad) % ld_unit != 0
        _geam_case(oracle, handle, dtype, 1, 0, m, n, 1.0, beta, 80 + m, lda_pad=lda_pad, ldb_pad=ldc_pad, ldc_pad=ldc_pad, inplace=True,
                   nan_b=(beta == 0.0))
        assert handle.last_kernel().startswith(cp_name[:-4] if "reg" in cp_name else cp_name), handle.last_kernel()
    # a row count that is not a multiple of 16 B on the all-TMA path: the bulk store clips at 16 B granularity, so the last
    # rows are stored by the transposers (this wrote row m of the padded target before)
    for (m, beta) in [(517, 0.0), (519, 0.5), (513, -1.0)]:
        pad = (-m) % 32
        _geam_case(oracle, handle, dtype, 1, 0, m, 300, 1.0, beta, 95 + m, ldb_pad=pad, ldc_pad=pad, inplace=True, nan_b=(beta == 0.0))
        assert handle.last_kernel().startswith(tma_name), handle.last_kernel()
    if "reg" in cp_name:
        return
    # the producers against the TMA loads on a source both can address
    m, n = 1000, 1100
    A = dev(oracle.fill(n * m, 5, dtype))
    outs = []
    try:
        for force in (0, 1):
            handle.set_option("geam.force_cpasync", force)
            Cm = torch.full((m * n,), float("nan"), dtype=torch.float64 if dtype == np.float64 else torch.float32, device="cuda")
            handle.geam(dtype == np.float64, 1, 0, m, n, 1.0, A, n, 0.0, Cm, m, Cm, m)
            outs.append((host(Cm), handle.last_kernel()))
    finally:
        handle.set_option("geam.force_cpasync", 0)
    assert outs[0][1] == "geam_transpose_tma" and outs[1][1] == "geam_transpose_cpasync_tma"
    assert np.array_equal(outs[0][0].view(np.uint8), outs[1][0].view(np.uint8))


def test_geam_argument_checks(handle):
    x = torch.zeros(64, dtype=torch.float64, device="cuda")
    y = torch.zeros(64, dtype=torch.float64, device="cuda")
    assert handle.geam(True, 0, 0, 8, 8, 1.0, x, 4, 0.0, y, 8, y, 8, check=False) == 7       # lda < m
    assert handle.geam(True, 1, 0, 8, 4, 1.0, x, 2, 0.0, y, 8, y, 8, check=False) == 7       # lda < n for op T
    assert handle.geam(True, 1, 0, 8, 8, 1.0, x, 8, 0.0, y, 8, x, 8, check=False) == 7       # in-place transpose
    assert handle.geam(True, 0, 0, 0, 8, 1.0, None, 1, 0.0, None, 1, None, 1, check=False) == 0


def test_stream_binding_and_introspection(handle):
    s = torch.cuda.Stream()
    h2 = rt.Handle(stream=s)
    assert h2.get_stream() == s.cuda_stream
    n0 = h2.launch_count()
    x = torch.ones(32 * 32, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    torch.cuda.synchronize()
    h2.geam(True, 1, 0, 32, 32, 1.0, x, 32, 0.0, y, 32, y, 32)
    s.synchronize()
    assert h2.launch_count() == n0 + 1 and h2.last_kernel().startswith("geam_transpose_reg")
    assert torch.equal(x, y)
    h2.close()


def test_stream_rebinding_orders_the_handles_scratch():
    """Two stream-K DGEMMs and a TRSM queued back to back on DIFFERENT streams through ONE handle share its scratch
    (stream-K slots and flags, TRSM inverses): rtat_b200_set_stream must order the new stream after the old one."""
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    h = rt.Handle(stream=s1)
    n = 1536  # 144 tiles of 128^2 on 148 SMs, 96 k-tiles: every CTA parks or merges stream-K partials
    A1, B1, A2, B2 = (torch.rand(n * n, dtype=torch.float64, device="cuda") - 0.5 for _ in range(4))
    C1, C2 = torch.empty_like(A1), torch.empty_like(A1)
    T = (torch.rand(n * n, dtype=torch.float64, device="cuda") - 0.5) / 64
    T.view(n, n).diagonal().add_(2.0)
    X = torch.rand(n * n, dtype=torch.float64, device="cuda")
    X0 = X.clone()
    torch.cuda.synchronize()
    for _ in range(3):
        X.copy_(X0)
        C1.fill_(float("nan"))
        C2.fill_(float("nan"))
        torch.cuda.synchronize()
        # no host synchronisation from here to the end of the round: only set_stream orders the four calls
        h.set_stream(s1)
        h.gemm(True, 0, 0, n, n, n, 1.0, A1, n, B1, n, 0.0, C1, n)
        assert "streamk" in h.last_kernel()
        h.set_stream(s2)
        h.trsm(True, 0, 0, 0, 0, n, n, 1.0, T, n, X, n)
        h.set_stream(s1)
        h.gemm(True, 0, 0, n, n, n, 1.0, A2, n, B2, n, 0.0, C2, n)
        h.set_stream(s2)
        torch.cuda.synchronize()
        col = lambda t: t.view(n, n).t()
        eps = np.finfo(np.float64).eps
        for Am, Bm, Cm in ((A1, B1, C1), (A2, B2, C2)):
            ref, bound = col(Am) @ col(Bm), col(Am).abs() @ col(Bm).abs()
            assert bool(torch.all((col(Cm) - ref).abs() <= 4 * n * eps * bound + eps * ref.abs()))
        L = torch.tril(col(T))
        res = L @ col(X) - col(X0)
        assert bool(torch.all(res.abs() <= 4 * n * eps * (L.abs() @ col(X).abs() + col(X0).abs())))
    h.close()


def test_fill_uniform_matches_oracle_generator(oracle, handle):
    for dtype, tdt in ((np.float64, torch.float64), (np.float32, torch.float32)):
        x = torch.empty(100003, dtype=tdt, device="cuda")
        handle.fill_uniform(x, 1234, -1.0, 1.0)
        assert np.array_equal(host(x), oracle.fill(100003, 1234, dtype))


# ---------------------------------------------------------------------------------------------
# BASELINE.json full-size configs: the reference's arithmetic at these sizes is cuBLAS, which
# torch.matmul reaches here (test-only use).  Both sides obey the stated bound against the exact
# product, so they may differ from each other by at most twice the bound.
# ---------------------------------------------------------------------------------------------
def _full_size_vs_cublas(handle, dtype, ta, tb, m, n, k, seed, streamk=1):
    tdt = torch.float64 if dtype == np.float64 else torch.float32
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    A = torch.empty(ar * ac, dtype=tdt, device="cuda")
    B = torch.empty(br * bc, dtype=tdt, device="cuda")
    handle.fill_uniform(A, seed)
    handle.fill_uniform(B, seed + 1)
    Cm = torch.full((m * n,), float("nan"), dtype=tdt, device="cuda")
    handle.set_option("dgemm.streamk", streamk)
    try:
        handle.gemm(dtype == np.float64, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, Cm, m)
    finally:
        handle.set_option("dgemm.streamk", 1)
    # column-major (rows, cols, ld=rows) buffer == row-major (cols, rows) tensor
    Am, Bm = A.view(ac, ar).t(), B.view(bc, br).t()
    opA, opB = (Am.t() if ta else Am), (Bm.t() if tb else Bm)
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = opA @ opB
        bound = opA.abs() @ opB.abs()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    got = Cm.view(n, m).t()
    eps = np.finfo(dtype).eps
    c = 2 if dtype == np.float64 else 8
    tol = 2 * c * k * eps * bound + eps * ref.abs()
    err = (got - ref).abs()
    name = handle.last_kernel()
    assert bool(torch.all(err <= tol)), f"max err {err.max().item():.3e} kernel={name}"
    # every entry against cuBLAS at twice the stated bound (both sides of the difference carry rounding error) — and a
    # sample of entries against the EXACT product at 1x the stated bound, |C - C_exact| <= c k eps (|A||B|) + eps |C_exact|
    ratio = sampled_exact_ratio(ORACLE(), dtype, ta, tb, m, n, k, A, ar, B, br, Cm, m)
    assert ratio <= 1.0, f"sampled |C - C_exact| is {ratio:.3g} x the stated bound, kernel={name}"
    return name


def test_dgemm_config1_1024_cubed(handle):
    assert "streamk" in _full_size_vs_cublas(handle, np.float64, 0, 0, 1024, 1024, 1024, 9001)
    _full_size_vs_cublas(handle, np.float64, 0, 0, 1024, 1024, 1024, 9001, streamk=0)


def test_dgemm_config3_unaligned(handle):
    # m=4097 n=3001 k=2049 op NT: odd leading dimensions => cp.async producers, ragged tiles, stream-K
    assert "cpasync8" in _full_size_vs_cublas(handle, np.float64, 0, 1, 4097, 3001, 2049, 9003)


def test_dgemm_config2_8192_tn(handle):
    assert "tma" in _full_size_vs_cublas(handle, np.float64, 1, 0, 8192, 8192, 8192, 9002)


def test_dgemm_config5_full_size_one_gpu(handle):
    """Config 5 itself, 32768^3 on one GPU (25.8 GB of operands, ~2 s): sampled entries against the exact product at 1x the
    stated bound; the full comparison against cuBLAS would cost a second 8.6 GB result, so only the sample is checked."""
    free_b, _ = torch.cuda.mem_get_info()
    if free_b < 30e9:
        pytest.skip("needs 30 GB of free device memory")
    n = 32768
    A = torch.empty(n * n, dtype=torch.float64, device="cuda")
    B = torch.empty(n * n, dtype=torch.float64, device="cuda")
    handle.fill_uniform(A, 9051)
    handle.fill_uniform(B, 9052)
    Cm = torch.full((n * n,), float("nan"), dtype=torch.float64, device="cuda")
    handle.gemm(True, 0, 0, n, n, n, 1.0, A, n, B, n, 0.0, Cm, n)
    torch.cuda.synchronize()
    assert "128x128" in handle.last_kernel(), handle.last_kernel()
    ratio = sampled_exact_ratio(ORACLE(), np.float64, 0, 0, n, n, n, A, n, B, n, Cm, n, samples=30)
    assert ratio <= 1.0, ratio
    assert not bool(torch.isnan(Cm[:: 4099]).any())   # a stride coprime to the tile sizes: nobody left NaN behind
    del A, B, Cm
    torch.cuda.empty_cache()


def test_dgemm_config5_column_block(handle):
    # one GPU's share of config 5 at 8 GPUs, scaled to fit the test budget: m=k=8192, n=1024
    _full_size_vs_cublas(handle, np.float64, 0, 0, 8192, 1024, 8192, 9005)


@pytest.mark.parametrize("tile,peel", [(64, 1), (128, 1), (128, 0), (192, 1), (192, 0)])
def test_dgemm_ragged_tiles_every_residue_class(handle, tile, peel):
    """The consumers of a ragged tile skip the DMMA blocks (64x64 tiles) or whole warps (128x128, 128x64 = tile 192) that lie
    outside C, and the ragged tiles are scheduled after the full ones: residues on either side of every 8 / 16 / 32 / 64
    boundary, both operand orientations (8-row blocks for a K-contiguous operand, pairs of blocks for an M/N-contiguous one),
    k long enough for the stream-K split (kt >= 8) and covering every k % 16 (the last k-tile skips the DMMA steps that see
    no valid k); C starts as NaN so an element nobody wrote is caught."""
    handle.set_option("dgemm.tile", tile)
    tile = 128 if tile == 192 else tile  # shapes below are sized in units of the larger tile edge
    handle.set_option("dgemm.peel", peel)  # 128x128 tiles: remainders of <= 16 rows / columns go to a second launch on 64x64 tiles
    try:
        res = [1, 7, 8, 9, 15, 16, 17, 31, 32, 33, 57, 63, 65, 100, 127]
        for i, rm in enumerate(res):
            rn = res[(i * 7 + 3) % len(res)]
            ta, tb = OPS[i % 4]
            _full_size_vs_cublas(handle, np.float64, ta, tb, 3 * tile + rm, 2 * tile + rn, 160 + i, 9200 + i)
        # many ragged tiles against few full ones, and the corner tile alone
        _full_size_vs_cublas(handle, np.float64, 0, 1, 20 * tile + 1, tile + 57, 2049, 9301)
        _full_size_vs_cublas(handle, np.float64, 1, 0, tile - 3, tile - 5, 300, 9302)
    finally:
        handle.set_option("dgemm.tile", 0)
        handle.set_option("dgemm.peel", 1)


@pytest.mark.parametrize("ta,tb", OPS)
@pytest.mark.parametrize("skinny", [1, 0], ids=["skinny", "strips64"])
def test_dgemm_thin_remainders_are_peeled_into_a_second_launch(oracle, handle, ta, tb, skinny):
    """m % 128 or n % 128 in 1..16 on 128x128 tiles (n % 64 on 128x64 tiles): the main launch covers the multiple of the tile,
    the row strip (all columns) and the column strip (main rows) go to the skinny kernel (dgemm.skinny = 0: to launches on
    64x64 tiles); every element written once, alpha / beta honoured in each piece, padding of C untouched, and the same
    result as the unpeeled kernel to within the bound.  The four op pairs put both strips through both skinny kernels
    (X contiguous along N / along K)."""
    handle.set_option("dgemm.tile", 128)
    handle.set_option("dgemm.skinny", skinny)

    def suffix(m, n, bn):  # strips up to 4 wide are the skinny kernel's, wider ones stay on 64x64 tiles
        thin = [x for x in (m % 128, n % bn) if x]
        return "+strips_skinny" if skinny and any(x <= 4 for x in thin) else "+strips64"

    try:
        for (m, n, k, beta) in [(641, 528, 200, 0.0), (1025, 643, 96, -0.5), (520, 1033, 130, 1.0), (513, 512, 64, 0.0),
                                (1040, 2049, 1111, 0.5), (2050, 517, 3000, 0.0)]:
            _gemm_case(oracle, handle, np.float64, ta, tb, m, n, k, 1.5, beta, 9500 + m, ldc_pad=3, nan_c=(beta == 0.0))
            assert handle.last_kernel().endswith(suffix(m, n, 128)), handle.last_kernel()
            handle.set_option("dgemm.peel", 0)
            _gemm_case(oracle, handle, np.float64, ta, tb, m, n, k, 1.5, beta, 9500 + m, ldc_pad=3, nan_c=(beta == 0.0))
            assert "strips" not in handle.last_kernel()
            handle.set_option("dgemm.peel", 1)
        # odd leading dimensions and element-aligned bases reach the strips too
        _gemm_case(oracle, handle, np.float64, ta, tb, 1025, 643, 333, -1.0, 0.25, 9590, lda_pad=1, ldb_pad=3, ldc_pad=1)
        assert handle.last_kernel().endswith(suffix(1025, 643, 128)), handle.last_kernel()
        # 128x64 tiles: m % 128 and n % 64 in 1..16 are the thin remainders
        handle.set_option("dgemm.tile", 192)
        for (m, n, k, beta) in [(641, 592, 200, 0.0), (1025, 579, 97, -0.5), (520, 1033, 130, 1.0), (640, 577, 73, 0.0)]:
            _gemm_case(oracle, handle, np.float64, ta, tb, m, n, k, 1.5, beta, 9600 + m, ldc_pad=3, nan_c=(beta == 0.0))
            assert "128x64" in handle.last_kernel() and handle.last_kernel().endswith(suffix(m, n, 64)), handle.last_kernel()
    finally:
        handle.set_option("dgemm.tile", 0)
        handle.set_option("dgemm.peel", 1)
        handle.set_option("dgemm.skinny", 1)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ta,tb", OPS)
def test_gemv_shaped_calls_take_the_skinny_kernel(oracle, handle, ta, tb, dtype):
    """C at most 4 rows or columns wide (and the other extent >= 256): the skinny FMA kernel (FP64, and FP32 with plain FP32
    arithmetic), both orientations of the large operand, K cut into chunks whose partial sums are added in order (several
    chunks / a single one), alpha / beta, odd leading dimensions, element-aligned bases; bit-identical when repeated."""
    dbl = dtype == np.float64
    prefix = ("d" if dbl else "s") + "gemm_skinny_fma"
    for (m, n, k, beta) in [(1, 3001, 2049, 0.0), (3, 700, 130, -0.5), (4, 260, 1000, 1.0), (2, 5000, 64, 0.0),
                            (3001, 1, 2049, 0.0), (700, 4, 130, 0.5), (300, 2, 4097, 0.0), (5000, 3, 17, 1.0)]:
        _gemm_case(oracle, handle, dtype, ta, tb, m, n, k, 1.5, beta, 9700 + m + n, ldc_pad=3, nan_c=(beta == 0.0))
        assert handle.last_kernel().startswith(prefix), handle.last_kernel()
        _gemm_case(oracle, handle, dtype, ta, tb, m, n, k, -1.0, beta, 9800 + m + n, lda_pad=1, ldb_pad=3, ldc_pad=1, offset=1)
        assert handle.last_kernel().startswith(prefix), handle.last_kernel()
    handle.set_option("dgemm.skinny", 0)
    try:
        _gemm_case(oracle, handle, dtype, ta, tb, 1, 3001, 200, 1.5, 0.0, 9790)
        assert "skinny" not in handle.last_kernel()
    finally:
        handle.set_option("dgemm.skinny", 1)
    m, n, k = (2, 4100, 3000)
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    tdt = torch.float64 if dbl else torch.float32
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.rand(ar * ac, dtype=tdt, device="cuda", generator=g) - 0.5
    B = torch.rand(br * bc, dtype=tdt, device="cuda", generator=g) - 0.5
    outs = []
    for _ in range(20):
        Cm = torch.zeros(m * n, dtype=tdt, device="cuda")
        handle.gemm(dbl, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, Cm, m)
        outs.append(Cm)
    assert handle.last_kernel().startswith(prefix)
    assert all(torch.equal(outs[0], o) for o in outs[1:])


def test_dgemm_l2_eviction_hints_do_not_change_results(handle):
    m, n, k = 1500, 1300, 700
    A = torch.empty(m * k, dtype=torch.float64, device="cuda")
    B = torch.empty(k * n, dtype=torch.float64, device="cuda")
    handle.fill_uniform(A, 5)
    handle.fill_uniform(B, 6)
    out = []
    try:
        for hints in (0, 1):
            handle.set_option("dgemm.l2_hints", hints)
            Cm = torch.full((m * n,), float("nan"), dtype=torch.float64, device="cuda")
            handle.gemm(True, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, Cm, m)
            assert "tma" in handle.last_kernel()
            out.append(Cm)
    finally:
        handle.set_option("dgemm.l2_hints", 0)
    assert torch.equal(out[0], out[1])


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gemm_extreme_aspect_ratios(handle, dtype):
    """Degenerate-looking shapes stress the schedules rather than the math: one tile with a very long k (every CTA of the
    stream-K grid contributes to a single tile), vectors, and single rows / columns of tiles."""
    for (ta, tb, m, n, k) in [(0, 0, 1, 1, 200000), (1, 0, 128, 128, 131072), (0, 1, 100000, 1, 3), (0, 0, 3, 100000, 2),
                              (1, 1, 1, 70000, 33), (0, 0, 40000, 130, 1)]:
        _full_size_vs_cublas(handle, dtype, ta, tb, m, n, k, 9100 + m % 97)


# ---------------------------------------------------------------------------------------------
# host-buffer entry points (what bench.py's e2e number times)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_gemm_host_entry(oracle, handle, dtype):
    for (ta, tb, m, n, k, beta, pad) in [(1, 0, 300, 200, 100, 0.0, 0), (0, 1, 129, 65, 33, -0.5, 3), (0, 0, 1000, 900, 700, 0.0, 0)]:
        ar, ac = (k, m) if ta else (m, k)
        br, bc = (n, k) if tb else (k, n)
        lda, ldb, ldc = ar + pad, br + pad, m + pad
        A, B, C0 = oracle.fill(lda * ac, 1, dtype), oracle.fill(ldb * bc, 2, dtype), oracle.fill(ldc * n, 3, dtype)
        hA, hB = torch.from_numpy(A).pin_memory(), torch.from_numpy(B)  # pinned and pageable
        hC = torch.from_numpy(C0.copy())
        for _ in range(2):  # second call reuses the staging arena
            hC.copy_(torch.from_numpy(C0))
            handle.gemm(dtype == np.float64, ta, tb, m, n, k, 1.0, hA, lda, hB, ldb, beta, hC, ldc, host=True)
            got = hC.numpy()
            ex, ab = oracle.gemm_exact(ta, tb, m, n, k, 1.0, A, lda, B, ldb, beta, C0, ldc)
            assert np.all(np.abs(view(got, m, n, ldc).astype(np.float64) - ex) <= gemm_tolerance(dtype, k, ab, ex))
            assert np.array_equal(view(got, ldc, n, ldc)[m:], view(C0, ldc, n, ldc)[m:])  # host padding rows untouched


@pytest.mark.parametrize("ta,tb", OPS)
def test_gemm_host_entry_k_panel_pipeline(handle, ta, tb):
    """A of 32 MiB and k >= 2048 switch the host entry to its two-phase schedule (K-panels of A against the leading
    column blocks, then whole-K column blocks): every op pair, beta != 0, padded leading dimensions.  Checked against
    cuBLAS' product (torch, test-only use) within twice the stated bound, like the full-size configs."""
    m, n, k, pad, alpha, beta = 2048, 600, 2051, 2, 0.75, -0.5
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    lda, ldb, ldc = ar + pad, br + pad, m + pad
    g = torch.Generator().manual_seed(77)
    hA = (torch.rand(lda * ac, dtype=torch.float64, generator=g) * 2 - 1).pin_memory()
    hB = torch.rand(ldb * bc, dtype=torch.float64, generator=g) * 2 - 1
    C0 = torch.rand(ldc * n, dtype=torch.float64, generator=g) * 2 - 1
    hC = C0.clone()
    launches = handle.launch_count()
    handle.gemm(True, ta, tb, m, n, k, alpha, hA, lda, hB, ldb, beta, hC, ldc, host=True)
    assert handle.launch_count() - launches >= 3, "expected one GEMM per K-panel under the leading blocks (3 panels here)"
    Am = hA.view(ac, lda).t()[:ar].cuda()
    Bm = hB.view(bc, ldb).t()[:br].cuda()
    opA, opB = (Am.t() if ta else Am), (Bm.t() if tb else Bm)
    Cin = C0.view(n, ldc).t()[:m].cuda()
    ref = alpha * (opA @ opB) + beta * Cin
    bound = abs(alpha) * (opA.abs() @ opB.abs()) + abs(beta) * Cin.abs()
    got = hC.view(n, ldc).t()
    eps = np.finfo(np.float64).eps
    assert bool(torch.all((got[:m].cuda() - ref).abs() <= 2 * 2 * k * eps * bound + 2 * eps * ref.abs()))
    assert torch.equal(got[m:], C0.view(n, ldc).t()[m:])  # host padding rows untouched


# ---------------------------------------------------------------------------------------------
# SGEMM on tcgen05 (3xTF32): forced with sgemm.variant = 2; needs 16 B aligned operands, ld % 4 == 0
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("ta,tb", OPS)
@pytest.mark.parametrize("presplit", [2, 1, 0])
def test_sgemm_tcgen05_3xtf32(oracle, handle, ta, tb, presplit):
    handle.set_option("sgemm.variant", 2)
    # 2: op(B) split once into the arena whatever the size; 1 (default): only from 768^3 multiply-adds on, or when TMA cannot
    # address B as stored; 0: split per tile in shared memory
    handle.set_option("sgemm.presplit", presplit)
    try:
        for (m, n, k) in [(128, 128, 32), (128, 128, 64), (256, 128, 96), (260, 132, 100), (64, 48, 40), (1, 1, 1), (513, 257, 33), (1024, 384, 520)]:
            # leading dimensions rounded up to a multiple of 4 floats (what TMA can address)
            ar = k if ta else m
            br = n if tb else k
            pad_a, pad_b = (-ar) % 4, (-br) % 4
            err = _gemm_case(oracle, handle, np.float32, ta, tb, m, n, k, 1.0, 0.0, 700 + m + k, lda_pad=pad_a, ldb_pad=pad_b)
            assert handle.last_kernel().startswith("sgemm_tcgen05"), handle.last_kernel()
            _gemm_case(oracle, handle, np.float32, ta, tb, m, n, k, -1.5, 0.75, 800 + m + k, lda_pad=pad_a + 4, ldb_pad=pad_b + 8, ldc_pad=3)
            _gemm_case(oracle, handle, np.float32, ta, tb, m, n, k, 1.0, 0.0, 900 + m + k, lda_pad=pad_a, ldb_pad=pad_b, nan_c=True)
            if presplit == 2 and m > 128:
                assert handle.last_kernel().endswith("presplitB"), handle.last_kernel()
            if presplit and m > 128:  # the pre-pass reads any ldb / alignment of B
                _gemm_case(oracle, handle, np.float32, ta, tb, m, n, k, 1.0, 0.0, 950 + m + k, lda_pad=pad_a, ldb_pad=pad_b + 1)
                assert handle.last_kernel().endswith("presplitB"), handle.last_kernel()
    finally:
        handle.set_option("sgemm.variant", 0)
        handle.set_option("sgemm.presplit", 1)


@pytest.mark.parametrize("ta,tb", OPS)
def test_sgemm_tcgen05_pads_an_op_a_tma_cannot_address(oracle, handle, ta, tb):
    """ld not a multiple of 4 floats / base 4 bytes off: op(A) is copied once into a 16 B-aligned buffer of the handle and
    the tensor-core kernel runs on the copy (before: the FP32 SIMT fallback); B goes through the pre-split pass."""
    for (m, n, k, lda_pad, offset) in [(300, 128, 96, 1, 0), (517, 257, 130, 3, 1), (1024, 64, 520, 2, 0), (129, 640, 33, 0, 1)]:
        ar = k if ta else m
        pad_a = (-ar) % 4 + lda_pad if lda_pad else (-ar) % 4
        _gemm_case(oracle, handle, np.float32, ta, tb, m, n, k, 1.25, -0.5, 1300 + m + k, lda_pad=pad_a, ldb_pad=1, ldc_pad=2, offset=offset)
        name = handle.last_kernel()
        unaligned = offset % 4 != 0 or (ar + pad_a) % 4 != 0
        assert name.startswith("sgemm_tcgen05") and name.endswith("_padA") == unaligned, name
    # too thin to pay for the extra pass over A: the FP32 kernel
    _gemm_case(oracle, handle, np.float32, ta, tb, 300, 48, 96, 1.0, 0.0, 1399, lda_pad=1)
    assert handle.last_kernel().startswith("sgemm_simt"), handle.last_kernel()


@pytest.mark.parametrize("ta,tb", OPS)
def test_sgemm_tcgen05_propagates_nan_from_either_operand(handle, ta, tb):
    """op(A) is split in registers with integer rounding (big) and an unrounded difference (small); a NaN of any payload —
    the canonical 0x7fffffff and 0xffffffff overflow the rounding add — must still reach every entry of its row of C, a NaN
    in op(B) every entry of its column, and nothing else."""
    m, n, k = 300, 200, 160
    for pattern in (0x7FC00000, 0x7FFFFFFF, 0xFFFFFFFF, 0x7F800001):
        for pair in (1, 0):
            handle.set_option("sgemm.variant", 2)
            handle.set_option("sgemm.pair", pair)
            try:
                ar, ac = (k, m) if ta else (m, k)
                br, bc = (n, k) if tb else (k, n)
                g = torch.Generator(device="cuda").manual_seed(pattern & 0xFFFF)
                A = torch.rand(ar * ac, dtype=torch.float32, device="cuda", generator=g) - 0.5
                B = torch.rand(br * bc, dtype=torch.float32, device="cuda", generator=g) - 0.5
                i0, l0, j0, l1 = 137, 45, 77, 101
                nan = torch.tensor([pattern - (1 << 32) if pattern >= (1 << 31) else pattern], dtype=torch.int32, device="cuda").view(torch.float32)
                A.view(ac, ar)[(i0, l0) if ta else (l0, i0)] = nan[0]       # op(A)[i0, l0]
                B.view(bc, br)[(l1, j0) if tb else (j0, l1)] = nan[0]       # op(B)[l1, j0]
                Cm = torch.zeros(m * n, dtype=torch.float32, device="cuda")
                handle.gemm(False, ta, tb, m, n, k, 1.0, A, ar, B, br, 0.0, Cm, m)
                torch.cuda.synchronize()
                got = torch.isnan(Cm.view(n, m).t())                          # (m, n)
                want = torch.zeros_like(got)
                want[i0, :] = True
                want[:, j0] = True
                assert torch.equal(got, want), (hex(pattern), pair, handle.last_kernel(), int(got.sum()), int(want.sum()))
            finally:
                handle.set_option("sgemm.variant", 0)
                handle.set_option("sgemm.pair", 1)


def test_sgemm_tcgen05_rejects_what_tma_cannot_address(handle):
    handle.set_option("sgemm.variant", 2)
    handle.set_option("sgemm.pad_a", 0)
    try:
        x = torch.zeros(70 * 70 + 8, dtype=torch.float32, device="cuda")
        assert handle.gemm(False, 0, 0, 70, 45, 62, 1.0, x, 70, x, 62, 0.0, x, 70, check=False) == 15  # ld = 70: not a multiple of 4
        assert handle.gemm(False, 0, 0, 64, 64, 64, 1.0, x.data_ptr() + 4, 64, x, 64, 0.0, x, 64, check=False) == 15  # 4 B aligned base
    finally:
        handle.set_option("sgemm.variant", 0)
        handle.set_option("sgemm.pad_a", 1)


def test_sgemm_accuracy_is_fp32_class_not_tf32_class(oracle, handle):
    """3xTF32 must be as accurate as an FP32 FMA chain: compare both against the exact product."""
    m, n, k = 256, 128, 4096
    A, B = oracle.fill(m * k, 11, np.float32), oracle.fill(k * n, 12, np.float32)
    ex, ab = oracle.gemm_exact(0, 0, m, n, k, 1.0, A, m, B, k, 0.0, None, m)
    errs = {}
    for variant in (1, 2):
        handle.set_option("sgemm.variant", variant)
        dC = torch.empty(m * n, dtype=torch.float32, device="cuda")
        handle.gemm(False, 0, 0, m, n, k, 1.0, dev(A), m, dev(B), k, 0.0, dC, m)
        errs[variant] = np.abs(view(host(dC), m, n, m).astype(np.float64) - ex).max()
    handle.set_option("sgemm.variant", 0)
    scale = np.abs(ex).max()
    assert errs[2] <= 4 * max(errs[1], 1e-7 * scale), errs      # same class as the FP32 kernel
    assert errs[2] <= 2e-5 * scale, errs                         # a single-pass TF32 product would be ~1e-3


def test_sgemm_config4_tall_skinny(handle):
    # config 4: m=131072 n=128 k=4096 op NN (2.1 GB of A); checked against cuBLAS FP32 via torch
    k = _full_size_vs_cublas(handle, np.float32, 0, 0, 131072, 128, 4096, 9004)
    assert k.startswith("sgemm_tcgen05"), k


def test_mg_panel_pipeline_on_one_gpu(oracle, handle):
    """rtat_b200_mg_dgemm with a distinct stage buffer: the K-panel copy/GEMM pipeline (here a local
    D2D copy stands in for the NVLink peer pull) must give the same C as one GEMM, for both op(A)."""
    for ta, tb in ((0, 0), (1, 0), (0, 1)):
        m, n, k = 300, 200, 530
        ar, ac = (k, m) if ta else (m, k)
        br, bc = (n, k) if tb else (k, n)
        lda = ar + 2
        A, B, C0 = oracle.fill(lda * ac, 1), oracle.fill(br * bc, 2), oracle.fill(m * n, 3)
        ex, ab = oracle.gemm_exact(ta, tb, m, n, k, 1.25, A, lda, B, br, -0.5, C0, m)
        cs = rt.ColumnSplit(handle, 0, 1)
        dA, dB = dev(A), dev(B)
        for panels in (0, 1, 5, 64):
            stage = torch.full_like(dA, float("nan"))
            for rep in range(2):  # second call reuses the stage: exercises the consumed-event handshake
                dC = dev(C0)
                cs.dgemm(ta, tb, m, n, k, 1.25, dA, lda, stage, dB, br, -0.5, dC, m, panels=panels)
                got = host(dC)
                assert np.all(np.abs(view(got, m, n, m) - ex) <= gemm_tolerance(np.float64, k, ab, ex)), (ta, tb, panels)
        # root path: no stage, nothing copied
        dC = dev(C0)
        cs.dgemm(ta, tb, m, n, k, 1.25, dA, lda, None, dB, br, -0.5, dC, m)
        assert np.all(np.abs(view(host(dC), m, n, m) - ex) <= gemm_tolerance(np.float64, k, ab, ex))
        cs.close()


def test_multi_gpu_host_entry_single_rank_and_all_visible_gpus():
    """rtat_b200_mg_dgemm_host: world 1 in-process path, then one rank per visible GPU (when the box has more than one)
    under torch.distributed.run; tests/mg_host_worker.py holds the checks."""
    import os
    import subprocess
    import sys

    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "mg_host_worker.py")
    env = {k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")}
    res = subprocess.run([sys.executable, worker], capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and "mg_host ok rank 0/1" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
    ngpu = torch.cuda.device_count()
    if ngpu >= 2:
        n = 2 if ngpu < 4 else 4
        res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
                              "--master-port", "29577", worker], capture_output=True, text=True, timeout=900, env=env)
        assert res.returncode == 0 and res.stdout.count("mg_host ok") == n, res.stdout[-2000:] + res.stderr[-4000:]
