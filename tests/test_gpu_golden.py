"""-m gpu: the CUDA path against the outputs of the UNMODIFIED reference on the same inputs (tests/golden/*, captured by
oracle/ref_driver.cpp from the reference's executors over cuBLAS 12.9 on a B200).  No oracle arithmetic sits between the
two sides here except the exact product that defines the GEMM bound:
  geam-type results (raw geam, MatrixMove, MatrixAccumulate)           bit-exact, padding included
  GEMM / SYRK plans and raw calls                                      both sides within c k eps (|A||B|) of the exact value
  everything a call must not touch (ld padding, the other triangle)    bit-exact
  TRSM plans and raw calls                                             relative difference <= 1e-12 (f64) / 1e-4 (f32)
Plans run through the same planner entry points a user calls (librtat_host.so), raw calls through the C ABI."""
import numpy as np
import pytest
import torch

import rtatblas_b200 as rt
from gpu_util import dev, host
from oracle_binding import view
from test_oracle_cpu import _golden as gemm_golden, golden_inputs, golden_output
from test_oracle_l3_cpu import _condition, _golden as l3_golden, _tri

pytestmark = pytest.mark.gpu


def _dt(case):
    return np.float64 if case["dtype"] == "f64" else np.float32


def _scratch(nbytes):
    return torch.full((nbytes + 16,), 0xFF, dtype=torch.uint8, device="cuda")  # NaN-patterned, like the golden run


def test_gemm_path_matches_reference_outputs(oracle, handle):
    cases, raw = gemm_golden()
    planners = {}
    seen = set()
    for case in cases:
        dt, kind = _dt(case), case["kind"]
        dbl = dt == np.float64
        eps = np.finfo(dt).eps
        gold = golden_output(case, raw)
        inp = golden_inputs(oracle, case)
        seen.add(kind)
        if kind in ("gemm_plan", "gemm_pad_plan", "raw_gemm"):
            ta, tb, m, n, k = case["transa"], case["transb"], case["m"], case["n"], case["k"]
            lda, ldb, ldc = inp["lda"], inp["ldb"], inp["ldc"]
            dA, dB, dC = dev(inp["A"]), dev(inp["B"]), dev(inp["C"])
            if kind == "raw_gemm":
                handle.gemm(dbl, ta, tb, m, n, k, case["alpha"], dA, lda, dB, ldb, case["beta"], dC, ldc)
            else:
                key = ("gemm" if kind == "gemm_plan" else "gemm_pad", dbl)
                if key not in planners:
                    planners[key] = rt.Planner(key[0], "double" if dbl else "float")
                pl = planners[key]
                need = pl.workspace(ta, tb, m, n, k, case["plan"])
                assert need == case["workspace_bytes"], case
                ws = _scratch(need)
                assert pl.gemm(handle, ta, tb, m, n, k, case["alpha"], dA, lda, dB, ldb, case["beta"], dC, ldc, ws, need, plan=case["plan"]) == case["plan"]
            got = host(dC)
            ex, ab = oracle.gemm_exact(ta, tb, m, n, k, case["alpha"], inp["A"], lda, inp["B"], ldb, case["beta"], inp["C"], ldc)
            tol = (2 if dbl else 8) * k * eps * ab + 2 * eps * np.abs(ex)
            assert np.all(np.abs(view(gold, m, n, ldc).astype(np.float64) - ex) <= tol), case
            assert np.all(np.abs(view(got, m, n, ldc).astype(np.float64) - ex) <= tol), case
            keep = np.ones(got.shape, bool)
            view(keep, m, n, ldc)[:] = False
            assert np.array_equal(got[keep].view(np.uint8), gold[keep].view(np.uint8)), case
        elif kind == "raw_geam":
            dA, dB, dC = dev(inp["A"]), dev(inp["B"]), dev(inp["C"])
            handle.geam(dbl, case["transa"], case["transb"], case["m"], case["n"], case["alpha"], dA, case["lda"], case["beta"], dB, case["ldb"], dC, case["ldc"])
            assert np.array_equal(host(dC).view(np.uint8), gold.view(np.uint8)), case
        elif kind == "move":
            dA, dB = dev(inp["A"]), dev(inp["B"])
            handle.geam(dbl, case["transpose"], 0, case["out_m"], case["out_n"], case["alpha"], dA, case["m"], 0.0, dB, case["out_ld"], dB, case["out_ld"])
            assert np.array_equal(host(dB).view(np.uint8), gold.view(np.uint8)), case
        elif kind == "accumulate":
            dA, dB = dev(inp["A"]), dev(inp["B"])
            handle.geam(dbl, case["transpose"], 0, case["m"], case["n"], case["alpha"], dA, case["lda"], case["beta"], dB, case["ldb"], dB, case["ldb"])
            assert np.array_equal(host(dB).view(np.uint8), gold.view(np.uint8)), case
    assert seen == {"gemm_plan", "gemm_pad_plan", "raw_gemm", "raw_geam", "move", "accumulate"}
    for pl in planners.values():
        pl.close()


def test_syrk_trsm_paths_match_reference_outputs(oracle, handle):
    cases, raw = l3_golden()
    planners = {}
    for case in cases:
        dt, kind, seed = _dt(case), case["kind"], case["seed"]
        dbl = dt == np.float64
        eps = np.finfo(dt).eps
        gold = np.frombuffer(raw, dtype=dt, count=case["count"], offset=case["offset"]).copy()
        if kind in ("syrk_plan", "raw_syrk"):
            lower, trans, n, k, alpha, beta = (case[x] for x in ("lower", "trans", "n", "k", "alpha", "beta"))
            ar, ac = (k, n) if trans else (n, k)
            lda, ldc = case.get("lda", ar), case.get("ldc", n)
            A, C0 = oracle.fill(lda * ac, seed, dt), oracle.fill(ldc * n, seed + 1, dt)
            dA, dC = dev(A), dev(C0)
            if kind == "raw_syrk":
                handle.syrk(dbl, 0 if lower else 1, trans, n, k, alpha, dA, lda, beta, dC, ldc)
            else:
                if ("syrk", dbl) not in planners:
                    planners[("syrk", dbl)] = rt.Planner("syrk", "double" if dbl else "float")
                pl = planners[("syrk", dbl)]
                need = pl.syrk_workspace(0 if lower else 1, trans, n, k, case["plan"])
                assert need == case["workspace_bytes"], case
                ws = _scratch(need)
                assert pl.syrk(handle, 0 if lower else 1, trans, n, k, alpha, dA, lda, beta, dC, ldc, ws, need, plan=case["plan"]) == case["plan"]
            got = host(dC)
            tri = _tri(n, lower)
            ex, ab = oracle.gemm_exact(trans, 1 - trans, n, n, k, alpha, A, lda, A, lda, beta, C0, ldc)
            tol = (2 if dbl else 8) * max(k, 1) * eps * ab + 2 * eps * np.abs(ex)
            assert np.all(np.abs(view(got, n, n, ldc).astype(np.float64) - ex)[tri] <= tol[tri]), case
            assert np.all(np.abs(view(gold, n, n, ldc).astype(np.float64) - ex)[tri] <= tol[tri]), case
            keep = np.ones(got.shape, bool)
            view(keep, n, n, ldc)[tri] = False
            # the other triangle (untouched, or beta-scaled by the transpose_C plans) and the ld padding: the reference's bits
            assert np.array_equal(got[keep].view(np.uint8), gold[keep].view(np.uint8)), case
        else:
            left, lower, trans, unit, m, n, alpha = (case[x] for x in ("left", "lower", "trans", "unit", "m", "n", "alpha"))
            na = m if left else n
            lda, ldb = case.get("lda", na), case.get("ldb", m)
            A = _condition(oracle.fill(lda * na, seed, dt), na, lda, unit)
            B0 = oracle.fill(ldb * n, seed + 1, dt)
            dA, dB = dev(A), dev(B0)
            args = (0 if left else 1, 0 if lower else 1, trans, unit, m, n)
            if kind == "raw_trsm":
                handle.trsm(dbl, *args, alpha, dA, lda, dB, ldb)
            else:
                if ("trsm", dbl) not in planners:
                    planners[("trsm", dbl)] = rt.Planner("trsm", "double" if dbl else "float")
                pl = planners[("trsm", dbl)]
                need = pl.trsm_workspace(*args, case["plan"])
                assert need == case["workspace_bytes"], case
                ws = _scratch(need)
                assert pl.trsm(handle, *args, alpha, dA, lda, dB, ldb, ws, need, plan=case["plan"]) == case["plan"]
            got = host(dB)
            G, W = view(gold, m, n, ldb).astype(np.float64), view(got, m, n, ldb).astype(np.float64)
            scale = max(np.abs(G).max(), 1.0)
            assert np.abs(G - W).max() <= (1e-12 if dbl else 1e-4) * scale, case
            keep = np.ones(got.shape, bool)
            view(keep, m, n, ldb)[:] = False
            assert np.array_equal(got[keep].view(np.uint8), gold[keep].view(np.uint8)), case
    for pl in planners.values():
        pl.close()
