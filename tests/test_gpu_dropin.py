"""-m gpu: the drop-in claim itself.  oracle/_ref/ref_driver_b200 and oracle/_ref/run_tests_b200 are the reference's
OWN, unmodified host sources (src/app/run_tests.cpp, src/methods/*.cpp, src/matrix_ops/*.h, src/planning/*.h,
src/timing/*.cpp) compiled with oracle/overlay/gpu-api.h first on the include path and linked against librtat_b200.so
instead of cuBLAS/cuRAND (oracle/Makefile).  So here the reference's GEMM_Executor / MatrixMove / MatrixAccumulate /
SYRK_Executor / TRSM_Executor run *their* code over *our* C ABI, and must reproduce what the same sources produced
over cuBLAS 12.9 (tests/golden/reference_b200*.bin):
  geam-type results, ld padding, untouched triangles      bit-exact
  GEMM / SYRK                                              within the stated bound of the exact product
  TRSM                                                     relative difference <= 1e-12 (f64) / 1e-4 (f32)
The binaries are prebuilt here (the GPU box has no /root/reference); the tests skip when they are absent."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle_binding import view
from test_oracle_cpu import _golden as gemm_golden, golden_inputs
from test_oracle_l3_cpu import _condition, _golden as l3_golden, _tri

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRIVER = os.path.join(ROOT, "oracle", "_ref", "ref_driver_b200")
RUN_TESTS = os.path.join(ROOT, "oracle", "_ref", "run_tests_b200")


def _no_cublas(path):
    out = subprocess.run(["ldd", path], capture_output=True, text=True).stdout
    assert "librtat_b200" in out and "cublas" not in out and "curand" not in out, out


def _run_driver(mode, tmp_path):
    if not os.path.exists(DRIVER):
        pytest.skip("oracle/_ref/ref_driver_b200 not built (needs /root/reference at build time)")
    _no_cublas(DRIVER)
    b, j = str(tmp_path / f"{mode}.bin"), str(tmp_path / f"{mode}.json")
    subprocess.run([DRIVER, mode, b, j], check=True, timeout=600)
    return json.load(open(j))["cases"], open(b, "rb").read()


def _arr(case, raw):
    dt = np.float64 if case["dtype"] == "f64" else np.float32
    return np.frombuffer(raw, dtype=dt, count=case["count"], offset=case["offset"]).copy()


def _same_bits(a, b):
    return np.array_equal(a.view(np.uint8), b.view(np.uint8))


def test_reference_gemm_path_over_our_abi_reproduces_its_cublas_outputs(oracle, tmp_path):
    cases, raw = _run_driver("golden", tmp_path)
    gcases, graw = gemm_golden()
    assert cases == gcases  # same case list, plans, workspace bytes, offsets
    kinds = set()
    for case in cases:
        got, gold = _arr(case, raw), _arr(case, graw)
        kinds.add(case["kind"])
        if case["kind"] in ("gemm_plan", "gemm_pad_plan", "raw_gemm"):
            dbl = case["dtype"] == "f64"
            eps = np.finfo(got.dtype).eps
            inp = golden_inputs(oracle, case)
            ta, tb, m, n, k = case["transa"], case["transb"], case["m"], case["n"], case["k"]
            ex, ab = oracle.gemm_exact(ta, tb, m, n, k, case["alpha"], inp["A"], inp["lda"], inp["B"], inp["ldb"], case["beta"], inp["C"], inp["ldc"])
            tol = (2 if dbl else 8) * k * eps * ab + 2 * eps * np.abs(ex)
            assert np.all(np.abs(view(got, m, n, inp["ldc"]).astype(np.float64) - ex) <= tol), case
            keep = np.ones(got.shape, bool)
            view(keep, m, n, inp["ldc"])[:] = False
            assert _same_bits(got[keep], gold[keep]), case
        else:  # raw_geam, move, accumulate
            assert _same_bits(got, gold), case
    assert kinds == {"gemm_plan", "gemm_pad_plan", "raw_gemm", "raw_geam", "move", "accumulate"}


def test_reference_syrk_trsm_paths_over_our_abi_reproduce_their_cublas_outputs(oracle, tmp_path):
    cases, raw = _run_driver("golden_l3", tmp_path)
    gcases, graw = l3_golden()
    assert cases == gcases
    for case in cases:
        got, gold = _arr(case, raw), _arr(case, graw)
        dt, seed = got.dtype.type, case["seed"]
        dbl = case["dtype"] == "f64"
        eps = np.finfo(dt).eps
        if case["kind"] in ("syrk_plan", "raw_syrk"):
            lower, trans, n, k, alpha, beta = (case[x] for x in ("lower", "trans", "n", "k", "alpha", "beta"))
            ar, ac = (k, n) if trans else (n, k)
            lda, ldc = case.get("lda", ar), case.get("ldc", n)
            A, C0 = oracle.fill(lda * ac, seed, dt), oracle.fill(ldc * n, seed + 1, dt)
            tri = _tri(n, lower)
            ex, ab = oracle.gemm_exact(trans, 1 - trans, n, n, k, alpha, A, lda, A, lda, beta, C0, ldc)
            tol = (2 if dbl else 8) * max(k, 1) * eps * ab + 2 * eps * np.abs(ex)
            assert np.all(np.abs(view(got, n, n, ldc).astype(np.float64) - ex)[tri] <= tol[tri]), case
            keep = np.ones(got.shape, bool)
            view(keep, n, n, ldc)[tri] = False
            assert _same_bits(got[keep], gold[keep]), case
        else:
            m, n = case["m"], case["n"]
            ldb = case.get("ldb", m)
            G, W = view(gold, m, n, ldb).astype(np.float64), view(got, m, n, ldb).astype(np.float64)
            assert np.abs(G - W).max() <= (1e-12 if dbl else 1e-4) * max(np.abs(G).max(), 1.0), case
            keep = np.ones(got.shape, bool)
            view(keep, m, n, ldb)[:] = False
            assert _same_bits(got[keep], gold[keep]), case


@pytest.mark.parametrize("method,dtype,nplans", [("gemm", "double", 8), ("gemm_pad", "double", 64), ("gemm", "float", 8),
                                                   ("syrk", "double", 4), ("trsm", "float", 4)])
def test_reference_run_tests_app_runs_on_our_abi(tmp_path, method, dtype, nplans):
    """The reference's benchmark driver (src/app/run_tests.cpp) itself, exhaustive and autotune, same JSON in/out."""
    if not os.path.exists(RUN_TESTS):
        pytest.skip("oracle/_ref/run_tests_b200 not built (needs /root/reference at build time)")
    _no_cublas(RUN_TESTS)
    if method in ("gemm", "gemm_pad"):
        problems = [{"transA": "T", "transB": "N", "m": 257, "n": 130, "k": 77}, {"transA": "N", "transB": "T", "m": 64, "n": 512, "k": 128}]
    elif method == "syrk":
        problems = [{"uplo": "Lower", "trans": "N", "n": 200, "k": 75}, {"uplo": "Upper", "trans": "T", "n": 129, "k": 300}]
    else:
        problems = [{"side": "Left", "uplo": "Lower", "trans": "N", "diag": "Non-Unit", "m": 300, "n": 77},
                    {"side": "Right", "uplo": "Upper", "trans": "T", "diag": "Unit", "m": 65, "n": 260}]
    for run_type in ("exhaustive", "autotune"):
        reps = 2 if run_type == "exhaustive" else nplans + 3
        spec = {"keywords": {"run_type": run_type, "method": method, "data_type": dtype, "repetitions": reps}, "problems": problems}
        p = tmp_path / f"{method}_{run_type}.json"
        p.write_text(json.dumps(spec))
        res = subprocess.run([RUN_TESTS, str(p)], capture_output=True, text=True, timeout=600)
        assert res.returncode == 0, res.stderr
        assert "GPU Error" not in res.stderr and "error" not in res.stderr.lower(), res.stderr
        out = json.loads(res.stdout[res.stdout.index("{"):])
        assert out["keywords"] == spec["keywords"]
        assert len(out["problems"]) == len(problems)
        for prob in out["problems"]:
            assert len(prob["options"]) == nplans
            counts = [len(o["times"]) for o in prob["options"]]
            if run_type == "exhaustive":
                assert counts == [reps] * nplans
            else:  # every plan sampled once, the remaining calls go to the converged plan
                assert sorted(counts) == [1] * (nplans - 1) + [4]
            assert all(t > 0 and np.isfinite(t) for o in prob["options"] for t in o["times"])
