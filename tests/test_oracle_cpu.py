"""CPU tests of the oracle (oracle/oracle.c): restated reference algorithm vs independent numpy
arithmetic, plan algebra (reference src/methods/gemm.cpp:82-134,182-221), workspace accounting
(src/matrix_ops/matrixop.h:164-189), and — when present — the golden outputs of the reference
executor itself (tests/golden/reference_b200.*)."""
import json
import os
from fractions import Fraction

import numpy as np
import pytest

from oracle_binding import GEMM_PAD_PLANS, GEMM_PLANS, view

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _inputs(orc, ta, tb, m, n, k, dtype, seed, lda=None, ldb=None, ldc=None):
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    lda, ldb, ldc = lda or ar, ldb or br, ldc or m
    A = orc.fill(lda * ac, seed, dtype)
    B = orc.fill(ldb * bc, seed + 1, dtype)
    Cm = orc.fill(ldc * n, seed + 2, dtype)
    return A, lda, B, ldb, Cm, ldc, (ar, ac, br, bc)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
@pytest.mark.parametrize("ta", [0, 1])
@pytest.mark.parametrize("tb", [0, 1])
def test_gemm_matches_numpy_within_the_stated_bound(oracle, dtype, ta, tb):
    m, n, k = 70, 45, 62  # reference tests/executor_test.cpp:15-17
    A, lda, B, ldb, Cm, ldc, (ar, ac, br, bc) = _inputs(oracle, ta, tb, m, n, k, dtype, 11, ldc=m + 3)
    C0 = Cm.copy()
    oracle.gemm(ta, tb, m, n, k, 1.5, A, lda, B, ldb, -0.5, Cm, ldc)
    Av, Bv = view(A, ar, ac, lda).astype(np.float64), view(B, br, bc, ldb).astype(np.float64)
    opA, opB = (Av.T if ta else Av), (Bv.T if tb else Bv)
    ref = 1.5 * opA @ opB - 0.5 * view(C0, m, n, ldc).astype(np.float64)
    bound = 1.5 * np.abs(opA) @ np.abs(opB) + 0.5 * np.abs(view(C0, m, n, ldc))
    eps = np.finfo(dtype).eps
    assert np.all(np.abs(view(Cm, m, n, ldc) - ref) <= 2 * k * eps * bound)
    # rows m..ldc of every column are padding and must be untouched
    assert np.array_equal(view(Cm, ldc, n, ldc)[m:], view(C0, ldc, n, ldc)[m:])
    ex, ab = oracle.gemm_exact(ta, tb, m, n, k, 1.5, A, lda, B, ldb, -0.5, C0, ldc)
    assert np.allclose(ex, ref, rtol=0, atol=4 * k * np.finfo(np.float64).eps * bound.max())
    assert np.allclose(ab, bound, rtol=1e-12)


def test_gemm_follows_reference_operation_order(oracle):
    # tests/common.h:185-190: C *= beta; then C += alpha*A*B for l ascending, all in T
    m, n, k = 5, 4, 9
    A, lda, B, ldb, Cm, ldc, _ = _inputs(oracle, 0, 0, m, n, k, np.float32, 3)
    C0 = Cm.copy()
    oracle.gemm(0, 0, m, n, k, np.float32(0.7), A, lda, B, ldb, np.float32(1.3), Cm, ldc)
    Av, Bv = view(A, m, k, lda), view(B, k, n, ldb)
    exp = np.empty((m, n), np.float32)
    for i in range(m):
        for j in range(n):
            c = np.float32(view(C0, m, n, ldc)[i, j] * np.float32(1.3))
            for l in range(k):
                c = np.float32(c + np.float32(np.float32(np.float32(0.7) * Av[i, l]) * Bv[l, j]))
            exp[i, j] = c
    assert np.array_equal(view(Cm, m, n, ldc), exp)


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_geam_definition_and_aliasing(oracle, dtype):
    m, n = 13, 7
    for ta in (0, 1):
        for tb in (0, 1):
            ar, ac = (n, m) if ta else (m, n)
            br, bc = (n, m) if tb else (m, n)
            A = oracle.fill((ar + 1) * ac, 5, dtype)
            B = oracle.fill((br + 2) * bc, 6, dtype)
            A[0] = -0.0
            Cm = np.full((m + 3) * n, 7.0, dtype)
            al, be = dtype(0.3), dtype(-1.7)
            oracle.geam(ta, tb, m, n, al, A, ar + 1, be, B, br + 2, Cm, m + 3)
            Av, Bv = view(A, ar, ac, ar + 1), view(B, br, bc, br + 2)
            opA, opB = (Av.T if ta else Av), (Bv.T if tb else Bv)
            pa = (al * opA).astype(dtype)
            if dtype == np.float32:  # float64 holds the product exactly => one rounding, i.e. an fma
                exp = (np.float64(be) * opB.astype(np.float64) + pa.astype(np.float64)).astype(dtype)
            else:  # exact rational arithmetic, correctly rounded once
                exp = np.array([[float(Fraction(float(be)) * Fraction(float(opB[i, j])) + Fraction(float(pa[i, j])))
                                 for j in range(n)] for i in range(m)])
            assert np.array_equal(view(Cm, m, n, m + 3), exp)
            assert np.all(view(Cm, m + 3, n, m + 3)[m:] == 7.0)
    # beta == 0: B is never read (may be NaN), alpha == 1 is a bit-exact copy incl. -0.0
    A = oracle.fill(m * n, 8, dtype)
    A[0] = -0.0
    Bn = np.full(m * n, np.nan, dtype)
    oracle.geam(0, 0, m, n, dtype(1), A, m, dtype(0), Bn, m, Bn, m)
    assert np.array_equal(Bn.view(np.uint8), A.view(np.uint8))
    # in place accumulate (reference matrixop.h:302): B = A^T + beta*B
    S = oracle.fill(n * m, 9, dtype)
    Bc = oracle.fill(m * n, 10, dtype)
    exp = view(S, n, m, n).T + (dtype(0.5) * view(Bc, m, n, m)).astype(dtype)  # 0.5*b is exact
    oracle.geam(1, 0, m, n, dtype(1), S, n, dtype(0.5), Bc, m, Bc, m)
    assert np.array_equal(view(Bc, m, n, m), exp.astype(dtype))


def _expected_workspace(plan, ta, tb, m, n, k, elem):
    """SURVEY §3.2 closed form, written independently of oracle.c."""
    pad = plan[3:] if len(plan) == 6 else "NNN"
    up = lambda x, p: -(-x // p) * p
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n, k) if tb else (k, n)
    ws = 0
    if plan[0] == "T" or pad[0] == "P":
        r, c = (ac, ar) if plan[0] == "T" else (ar, ac)
        ws += up(r, 32 if pad[0] == "P" else 1) * c
    if plan[1] == "T" or pad[1] == "P":
        r, c = (bc, br) if plan[1] == "T" else (br, bc)
        ws += up(r, 32 if pad[1] == "P" else 1) * c
    if plan[2] == "T":
        ws += up(n, 32 if pad[2] == "P" else 1) * m
    elif pad[2] == "P":
        ws += up(m, 32) * n
    return ws * elem


@pytest.mark.parametrize("plans", [GEMM_PLANS, GEMM_PAD_PLANS])
def test_plan_workspace_accounting(oracle, plans):
    for ta in (0, 1):
        for tb in (0, 1):
            for plan in plans:
                for (m, n, k) in [(70, 45, 62), (4097, 3001, 2049), (33, 64, 1)]:
                    assert oracle.plan_workspace(plan, ta, tb, m, n, k, 8) == _expected_workspace(plan, ta, tb, m, n, k, 8)
    assert oracle.plan_workspace("NNN", 0, 0, 10, 10, 10, 8) == 0
    assert oracle.plan_workspace("XYZ", 0, 0, 1, 1, 1, 8) == 2**64 - 1
    assert oracle.plan_workspace("NNNN", 0, 0, 1, 1, 1, 8) == 2**64 - 1
    # config 2 (TN 8192^3): explicit-transpose plan TNN needs one 512 MiB scratch, TTT 1.5 GiB
    assert oracle.plan_workspace("TNN", 1, 0, 8192, 8192, 8192, 8) == 512 << 20
    assert oracle.plan_workspace("TTT", 1, 0, 8192, 8192, 8192, 8) == 3 * (512 << 20)


@pytest.mark.parametrize("dtype,plans", [(np.float64, GEMM_PLANS), (np.float64, GEMM_PAD_PLANS), (np.float32, GEMM_PLANS)])
def test_every_plan_is_algebraically_the_same_gemm(oracle, dtype, plans):
    m, n, k = 23, 16, 35  # reference tests/planning_test.cpp:89-91
    eps = np.finfo(dtype).eps
    for ta in (0, 1):
        for tb in (0, 1):
            A, lda, B, ldb, Cm, ldc, _ = _inputs(oracle, ta, tb, m, n, k, dtype, 21 + 2 * ta + tb)
            ex, ab = oracle.gemm_exact(ta, tb, m, n, k, 1.0, A, lda, B, ldb, 0.5, Cm, ldc)
            for plan in plans:
                Cp = Cm.copy()
                rc, ws = oracle.gemm_plan(plan, ta, tb, m, n, k, dtype(1.0), A, lda, B, ldb, dtype(0.5), Cp, ldc)
                assert rc == 0
                assert np.all(np.abs(view(Cp, m, n, ldc) - ex) <= 2 * k * eps * ab + eps * np.abs(ex)), plan
                assert np.isnan(ws[-1])  # one-past-the-end sentinel untouched: layout fits the accounted bytes


def test_plan_with_too_little_workspace_is_rejected(oracle):
    # reference src/methods/executor.h:49-51
    m, n, k = 8, 8, 8
    A, lda, B, ldb, Cm, ldc, _ = _inputs(oracle, 0, 0, m, n, k, np.float64, 1)
    rc, _ = oracle.gemm_plan("TNN", 0, 0, m, n, k, 1.0, A, lda, B, ldb, 0.0, Cm, ldc, ws=np.empty(4))
    assert rc == 2
    rc, _ = oracle.gemm_plan("QNN", 0, 0, m, n, k, 1.0, A, lda, B, ldb, 0.0, Cm, ldc, ws=np.empty(4))
    assert rc == 1


def test_fill_is_counter_based_and_in_range(oracle):
    a = oracle.fill(1000, 42)
    b = oracle.fill(10, 42)
    assert np.array_equal(a[:10], b)
    assert a.min() > -1.0 and a.max() <= 1.0 and abs(a.mean()) < 0.1
    f = oracle.fill(1000, 42, np.float32)
    assert np.array_equal(f, a.astype(np.float32))


# ---- pinning against the reference executor's own outputs -------------------------------------
def _golden():
    jp, bp = os.path.join(GOLD, "reference_b200.json"), os.path.join(GOLD, "reference_b200.bin")
    if not (os.path.exists(jp) and os.path.exists(bp)):
        pytest.skip("golden vectors not generated yet (tests/golden/make_golden.sh on a GPU box)")
    meta = json.load(open(jp))
    raw = open(bp, "rb").read()
    return meta["cases"], raw


def golden_output(case, raw):
    dt = np.float64 if case["dtype"] == "f64" else np.float32
    return np.frombuffer(raw, dtype=dt, count=case["count"], offset=case["offset"]).copy()


def golden_inputs(orc, case):
    """Regenerate the seeded inputs of a golden case exactly as oracle/ref_driver.cpp does."""
    dt = np.float64 if case["dtype"] == "f64" else np.float32
    kind, seed = case["kind"], case["seed"]

    def specials(v):
        v[0] = -0.0
        v[1] = 1e-310 if dt == np.float64 else np.float32(1e-42)
        v[2] = 1e300 if dt == np.float64 else np.float32(1e30)
        v[3] = 0.0
        return v

    if kind in ("gemm_plan", "gemm_pad_plan", "raw_gemm"):
        ta, tb, m, n, k = case["transa"], case["transb"], case["m"], case["n"], case["k"]
        ar, ac = (k, m) if ta else (m, k)
        br, bc = (n, k) if tb else (k, n)
        lda, ldb, ldc = case.get("lda", ar), case.get("ldb", br), case.get("ldc", m)
        return dict(A=orc.fill(lda * ac, seed, dt), lda=lda, B=orc.fill(ldb * bc, seed + 1, dt), ldb=ldb,
                    C=orc.fill(ldc * n, seed + 2, dt), ldc=ldc)
    if kind == "raw_geam":
        ta, tb, m, n = case["transa"], case["transb"], case["m"], case["n"]
        ac = m if ta else n
        bc = m if tb else n
        return dict(A=specials(orc.fill(case["lda"] * ac, seed, dt)), B=orc.fill(case["ldb"] * bc, seed + 1, dt),
                    C=orc.fill(case["ldc"] * n, seed + 2, dt))
    if kind == "move":
        return dict(A=specials(orc.fill(case["m"] * case["n"], seed, dt)), B=orc.fill(case["out_ld"] * case["out_n"], seed + 1, dt))
    if kind == "accumulate":
        m, n = case["m"], case["n"]
        ac = m if case["transpose"] else n
        return dict(A=specials(orc.fill(case["lda"] * ac, seed, dt)), B=orc.fill(case["ldb"] * n, seed + 1, dt))
    raise AssertionError(kind)


def test_oracle_matches_reference_executor_goldens(oracle):
    cases, raw = _golden()
    seen = set()
    for case in cases:
        dt = np.float64 if case["dtype"] == "f64" else np.float32
        eps = np.finfo(dt).eps
        gold = golden_output(case, raw)
        inp = golden_inputs(oracle, case)
        kind = case["kind"]
        seen.add(kind)
        if kind in ("gemm_plan", "gemm_pad_plan"):
            ta, tb, m, n, k = case["transa"], case["transb"], case["m"], case["n"], case["k"]
            assert oracle.plan_workspace(case["plan"], ta, tb, m, n, k, dt().itemsize) == case["workspace_bytes"], case
            Cp = inp["C"].copy()
            rc, _ = oracle.gemm_plan(case["plan"], ta, tb, m, n, k, dt(case["alpha"]), inp["A"], inp["lda"], inp["B"], inp["ldb"], dt(case["beta"]), Cp, inp["ldc"])
            assert rc == 0
            ex, ab = oracle.gemm_exact(ta, tb, m, n, k, case["alpha"], inp["A"], inp["lda"], inp["B"], inp["ldb"], case["beta"], inp["C"], inp["ldc"])
            c = 2 if dt == np.float64 else 8
            tol = c * k * eps * ab + eps * np.abs(ex)
            assert np.all(np.abs(view(gold, m, n, m) - ex) <= tol), case
            assert np.all(np.abs(view(Cp, m, n, m) - ex) <= tol), case
        elif kind == "raw_gemm":
            ta, tb, m, n, k = case["transa"], case["transb"], case["m"], case["n"], case["k"]
            Cp = inp["C"].copy()
            oracle.gemm(ta, tb, m, n, k, dt(case["alpha"]), inp["A"], inp["lda"], inp["B"], inp["ldb"], dt(case["beta"]), Cp, inp["ldc"])
            ex, ab = oracle.gemm_exact(ta, tb, m, n, k, case["alpha"], inp["A"], inp["lda"], inp["B"], inp["ldb"], case["beta"], inp["C"], inp["ldc"])
            c = 2 if dt == np.float64 else 8
            tol = c * k * eps * ab + eps * np.abs(ex)
            ldc = inp["ldc"]
            assert np.all(np.abs(view(gold, m, n, ldc) - ex) <= tol), case
            assert np.all(np.abs(view(Cp, m, n, ldc) - ex) <= tol), case
            # padding rows untouched by the reference, too
            assert np.array_equal(view(gold, ldc, n, ldc)[m:], view(inp["C"], ldc, n, ldc)[m:])
        elif kind == "raw_geam":
            Cp = inp["C"].copy()
            oracle.geam(case["transa"], case["transb"], case["m"], case["n"], dt(case["alpha"]), inp["A"], case["lda"], dt(case["beta"]), inp["B"], case["ldb"], Cp, case["ldc"])
            assert np.array_equal(Cp.view(np.uint8), gold.view(np.uint8)), case  # bit-exact incl. padding
        elif kind == "move":
            Bp = inp["B"].copy()
            oracle.geam(case["transpose"], 0, case["out_m"], case["out_n"], dt(case["alpha"]), inp["A"], case["m"], dt(0), Bp, case["out_ld"], Bp, case["out_ld"])
            assert np.array_equal(Bp.view(np.uint8), gold.view(np.uint8)), case
        elif kind == "accumulate":
            Bp = inp["B"].copy()
            oracle.geam(case["transpose"], 0, case["m"], case["n"], dt(case["alpha"]), inp["A"], case["lda"], dt(case["beta"]), Bp, case["ldb"], Bp, case["ldb"])
            assert np.array_equal(Bp.view(np.uint8), gold.view(np.uint8)), case
    assert seen == {"gemm_plan", "gemm_pad_plan", "raw_gemm", "raw_geam", "move", "accumulate"}
