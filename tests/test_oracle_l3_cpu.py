"""CPU tests of the oracle's SYRK / TRSM restatement (oracle/oracle.h) and its pinning against the outputs of the
UNMODIFIED reference executors (SYRK_Executor / TRSM_Executor over cuBLAS 12.9, captured on a B200 by
oracle/ref_driver.cpp golden_l3 with NaN-patterned scratch; tests/golden/reference_b200_l3.*)."""
import json
import os

import numpy as np
import pytest

from oracle_binding import BOOL_PLANS, view

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _tri(n, lower):
    return np.tril(np.ones((n, n), bool)) if lower else np.triu(np.ones((n, n), bool))


def _condition(A, na, lda, unit):
    """condition_triangle of oracle/ref_driver.cpp: off-diagonals / 4, non-unit diagonal pushed away from zero by 2."""
    V = view(A, na, na, lda)
    d = np.arange(na)
    diag = V[d, d].copy()
    V[:] = V / A.dtype.type(4)
    V[d, d] = diag if unit else diag + np.where(diag < 0, -2, 2).astype(A.dtype)
    return A


def _solve_exact(A, na, lda, left, lower, trans, unit, rhs):
    T = (np.tril(view(A, na, na, lda)) if lower else np.triu(view(A, na, na, lda))).astype(np.float64)
    if unit:
        np.fill_diagonal(T, 1.0)
    op = T.T if trans else T
    return (np.linalg.solve(op, rhs) if left else np.linalg.solve(op.T, rhs.T).T), op


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_syrk_oracle_is_the_gemm_triangle_and_plans_agree(oracle, dtype):
    n, k, alpha, beta = 37, 19, 0.7, -0.5
    for lower in (0, 1):
        for trans in (0, 1):
            ar, ac = (k, n) if trans else (n, k)
            A, C0 = oracle.fill(ar * ac, 1, dtype), oracle.fill(n * n, 2, dtype)
            full = C0.copy()
            oracle.gemm(trans, 1 - trans, n, n, k, alpha, A, ar, A, ar, beta, full, n)
            tri = _tri(n, lower)
            Cs = C0.copy()
            oracle.syrk(lower, trans, n, k, alpha, A, ar, beta, Cs, n)
            # same loop, same order: bit-identical to the GEMM oracle on the triangle, untouched elsewhere
            assert np.array_equal(view(Cs, n, n, n)[tri], view(full, n, n, n)[tri])
            assert np.array_equal(view(Cs, n, n, n)[~tri], view(C0, n, n, n)[~tri])
            for plan in BOOL_PLANS:
                Cp = C0.copy()
                rc, ws = oracle.syrk_plan(plan, lower, trans, n, k, alpha, A, ar, beta, Cp, n)
                assert rc == 0
                assert oracle.syrk_plan_workspace(plan, trans, n, k, A.itemsize) == A.itemsize * ((n * k if plan[0] == "T" else 0) + (n * n if plan[1] == "T" else 0))
                V = view(Cp, n, n, n)
                assert np.allclose(V[tri], view(full, n, n, n)[tri], rtol=0, atol=64 * np.finfo(dtype).eps)
                # transpose_C plans scale the triangle BLAS leaves alone by beta (syrk.cpp:92-97); the others keep it
                want_other = view(C0, n, n, n)[~tri] * dtype(beta) if plan[1] == "T" else view(C0, n, n, n)[~tri]
                assert np.array_equal(V[~tri], want_other.astype(dtype))
    assert oracle.syrk_plan_workspace("NN", 0, 4, 4, 8) == 2**64 - 1  # not Bool_Op letters
    Cp = oracle.fill(16, 3, dtype)
    ws = np.zeros(1, dtype)
    assert oracle.lib.oracle_dsyrk_plan(b"TT", 0, 0, 4, 4, 1.0, Cp.ctypes.data, 4, 0.0, Cp.ctypes.data, 4, ws.ctypes.data, 8) == 2  # workspace too small


@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_trsm_oracle_solves_and_plans_agree(oracle, dtype):
    m, n, alpha = 21, 17, 1.5
    tol = 1e-12 if dtype == np.float64 else 2e-5
    for left in (0, 1):
        na = m if left else n
        for lower in (0, 1):
            for trans in (0, 1):
                for unit in (0, 1):
                    A = _condition(oracle.fill(na * na, 3, dtype), na, na, unit)
                    B0 = oracle.fill(m * n, 4, dtype)
                    X, _ = _solve_exact(A, na, na, left, lower, trans, unit, alpha * view(B0, m, n, m).astype(np.float64))
                    # what must not be read is poisoned: other triangle, and the diagonal when unit
                    Ap = A.copy()
                    V = view(Ap, na, na, na)
                    V[~_tri(na, lower)] = np.nan
                    if unit:
                        V[np.arange(na), np.arange(na)] = np.nan
                    B = B0.copy()
                    oracle.trsm(left, lower, trans, unit, m, n, alpha, Ap, na, B, m)
                    assert np.abs(view(B, m, n, m) - X).max() < tol * max(1.0, np.abs(X).max())
                    for plan in BOOL_PLANS:
                        Bp = B0.copy()
                        rc, _ = oracle.trsm_plan(plan, left, lower, trans, unit, m, n, alpha, A, na, Bp, m)
                        assert rc == 0
                        assert oracle.trsm_plan_workspace(plan, left, m, n, A.itemsize) == A.itemsize * ((m * n if plan[0] == "T" else 0) + (na * na if plan[1] == "T" else 0))
                        assert np.abs(view(Bp, m, n, m) - X).max() < tol * max(1.0, np.abs(X).max()), (plan, left, lower, trans, unit)


def _golden():
    jp, bp = os.path.join(GOLD, "reference_b200_l3.json"), os.path.join(GOLD, "reference_b200_l3.bin")
    meta = json.load(open(jp))
    return meta["cases"], open(bp, "rb").read()


def test_oracle_matches_reference_syrk_trsm_goldens(oracle):
    cases, raw = _golden()
    seen = set()
    for case in cases:
        dt = np.float64 if case["dtype"] == "f64" else np.float32
        eps = np.finfo(dt).eps
        gold = np.frombuffer(raw, dtype=dt, count=case["count"], offset=case["offset"]).copy()
        kind, seed = case["kind"], case["seed"]
        seen.add(kind)
        if kind in ("syrk_plan", "raw_syrk"):
            lower, trans, n, k, alpha, beta = case["lower"], case["trans"], case["n"], case["k"], case["alpha"], case["beta"]
            ar, ac = (k, n) if trans else (n, k)
            lda, ldc = case.get("lda", ar), case.get("ldc", n)
            A, C0 = oracle.fill(lda * ac, seed, dt), oracle.fill(ldc * n, seed + 1, dt)
            want = C0.copy()
            if kind == "syrk_plan":
                assert case["key"] == f"{'Lower' if lower else 'Upper'},{'T' if trans else 'N'},{n},{k}"
                assert oracle.syrk_plan_workspace(case["plan"], trans, n, k, dt().itemsize) == case["workspace_bytes"], case
                rc, _ = oracle.syrk_plan(case["plan"], lower, trans, n, k, dt(alpha), A, lda, dt(beta), want, ldc)
                assert rc == 0
            else:
                oracle.syrk(lower, trans, n, k, dt(alpha), A, lda, dt(beta), want, ldc)
            tri = _tri(n, lower)
            ex, ab = oracle.gemm_exact(trans, 1 - trans, n, n, k, alpha, A, lda, A, lda, beta, C0, ldc)
            c = 2 if dt == np.float64 else 8
            tol = c * max(k, 1) * eps * ab + 2 * eps * np.abs(ex)
            G, W = view(gold, n, n, ldc), view(want, n, n, ldc)
            # owned triangle: the reference's cuBLAS result and the oracle both within the GEMM bound of the exact value
            assert np.all(np.abs(G.astype(np.float64) - ex)[tri] <= tol[tri]), case
            assert np.all(np.abs(W.astype(np.float64) - ex)[tri] <= tol[tri]), case
            # everything else (other triangle incl. the transpose_C plans' beta scaling, ld padding): bit-exact
            keep = np.ones(gold.shape, bool)
            view(keep, n, n, ldc)[tri] = False
            assert np.array_equal(gold[keep].view(np.uint8), want[keep].view(np.uint8)), case
        elif kind in ("trsm_plan", "raw_trsm"):
            left, lower, trans, unit, m, n, alpha = (case[x] for x in ("left", "lower", "trans", "unit", "m", "n", "alpha"))
            na = m if left else n
            lda, ldb = case.get("lda", na), case.get("ldb", m)
            A = _condition(oracle.fill(lda * na, seed, dt), na, lda, unit)
            B0 = oracle.fill(ldb * n, seed + 1, dt)
            want = B0.copy()
            if kind == "trsm_plan":
                assert case["key"] == f"{'Left' if left else 'Right'},{'Lower' if lower else 'Upper'},{'T' if trans else 'N'},{'Unit' if unit else 'Non-Unit'},{m},{n}"
                assert oracle.trsm_plan_workspace(case["plan"], left, m, n, dt().itemsize) == case["workspace_bytes"], case
                rc, _ = oracle.trsm_plan(case["plan"], left, lower, trans, unit, m, n, dt(alpha), A, lda, want, ldb)
                assert rc == 0
            else:
                oracle.trsm(left, lower, trans, unit, m, n, dt(alpha), A, lda, want, ldb)
            rhs = alpha * view(B0, m, n, ldb).astype(np.float64)
            X, op = _solve_exact(A, na, lda, left, lower, trans, unit, rhs)
            G, W = view(gold, m, n, ldb).astype(np.float64), view(want, m, n, ldb).astype(np.float64)
            for sol, who in ((G, "reference"), (W, "oracle")):
                res = (op @ sol - rhs) if left else (sol @ op - rhs)
                bound = (np.abs(op) @ np.abs(sol) + np.abs(rhs)) if left else (np.abs(sol) @ np.abs(op) + np.abs(rhs))
                assert np.all(np.abs(res) <= 4 * na * eps * bound), (who, case)
            scale = max(np.abs(X).max(), 1.0)
            assert np.abs(G - W).max() <= (1e-12 if dt == np.float64 else 1e-4) * scale, case
            keep = np.ones(gold.shape, bool)
            view(keep, m, n, ldb)[:] = False
            assert np.array_equal(gold[keep].view(np.uint8), want[keep].view(np.uint8)), case
        else:
            raise AssertionError(f"unknown golden kind {kind}")
    assert seen == {"syrk_plan", "raw_syrk", "trsm_plan", "raw_trsm"}
    assert len(cases) == 240
