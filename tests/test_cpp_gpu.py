"""-m gpu: the restated reference GoogleTest suites (tests/cpp/gpu_tests.cpp) and the run_tests
driver on the B200 back end, plus the planner driven from Python against the oracle."""
import json
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_reference_suites_restated_in_cpp():
    r = subprocess.run([os.path.join(ROOT, "build", "bin", "gpu_tests")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-6000:] + r.stderr[-3000:]
    assert " 0 failures" in r.stdout and r.stdout.count("[  OK  ]") >= 14


@pytest.mark.parametrize("method,dtype,run_type,nplans", [("gemm", "double", "exhaustive", 8), ("gemm_pad", "double", "exhaustive", 64),
                                                         ("gemm", "float", "autotune", 8)])
def test_run_tests_driver_schema(tmp_path, method, dtype, run_type, nplans):
    """Same input/output documents as the reference's run_tests (src/app/run_tests.cpp:55-72)."""
    doc = {"keywords": {"run_type": run_type, "method": method, "data_type": dtype, "repetitions": 12 if run_type == "autotune" else 2},
           "problems": [{"transA": "N", "transB": "T", "m": 257, "n": 129, "k": 65}, {"transA": "T", "transB": "N", "m": 64, "n": 64, "k": 64}]}
    f = tmp_path / "in.json"
    f.write_text(json.dumps(doc))
    r = subprocess.run([os.path.join(ROOT, "build", "bin", "run_tests"), str(f)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads(r.stdout)
    assert out["keywords"] == doc["keywords"]
    assert len(out["problems"]) == 2
    # map order: key text "N,T,,257,65,129" < "T,N,,64,64,64"
    assert out["problems"][0]["key"] == {"transA": "N", "transB": "T", "m": 257, "n": 129, "k": 65}
    for prob in out["problems"]:
        assert len(prob["options"]) == nplans
        names = ["".join(o["option"][k] for k in (["transA", "transB", "transC"] + (["padA", "padB", "padC"] if nplans == 64 else []))) for o in prob["options"]]
        assert names == sorted(names)
        total = sum(len(o["times"]) for o in prob["options"])
        if run_type == "exhaustive":
            assert all(len(o["times"]) == 2 and all(t > 0 for t in o["times"]) for o in prob["options"])
        else:
            assert total == 12 and all(len(o["times"]) >= 1 for o in prob["options"])


def test_planner_every_plan_matches_oracle_plan_algebra(oracle, handle):
    """Each plan through the C++ executor on the GPU vs the oracle's CPU execution of the same
    operation tree, including the scratch contents the explicit moves leave behind (bit-exact)."""
    import torch
    import rtatblas_b200 as rt
    from gpu_util import dev, gemm_tolerance, host
    from oracle_binding import GEMM_PAD_PLANS, GEMM_PLANS, view

    for method, plans, dt, np_dt in (("gemm", GEMM_PLANS, "double", np.float64), ("gemm_pad", GEMM_PAD_PLANS, "double", np.float64),
                                     ("gemm", GEMM_PLANS, "float", np.float32)):
        planner = rt.Planner(method, dt)
        for ta, tb in ((0, 1), (1, 0)):
            m, n, k = 37, 21, 19
            ar, ac = (k, m) if ta else (m, k)
            br, bc = (n, k) if tb else (k, n)
            A, B, C0 = oracle.fill(ar * ac, 1, np_dt), oracle.fill(br * bc, 2, np_dt), oracle.fill(m * n, 3, np_dt)
            ex, ab = oracle.gemm_exact(ta, tb, m, n, k, 1.5, A, ar, B, br, -0.5, C0, m)
            tol = gemm_tolerance(np_dt, k, ab, ex)
            for plan in plans:
                need = planner.workspace(ta, tb, m, n, k, plan)
                ws = torch.full((need // A.itemsize + 1,), float("nan"), dtype=torch.float64 if np_dt == np.float64 else torch.float32, device="cuda")
                dA, dB, dC = dev(A), dev(B), dev(C0)
                used = planner.gemm(handle, ta, tb, m, n, k, 1.5, dA, ar, dB, br, -0.5, dC, m, ws, need, plan=plan)
                assert used == plan
                got = host(dC)
                assert np.all(np.abs(view(got, m, n, m).astype(np.float64) - ex) <= tol), (method, plan)
                # scratch: one-past-the-end sentinel intact, explicit A'/B' copies bit-identical to the oracle's
                Cp = C0.copy()
                rc, ows = oracle.gemm_plan(plan, ta, tb, m, n, k, np_dt(1.5), A, ar, B, br, np_dt(-0.5), Cp, m)
                assert rc == 0
                gws = host(ws)
                assert np.isnan(gws[-1])
                if plan[2] == "N" and (len(plan) == 3 or plan[5] == "N"):  # no S in front: scratch is [A'][B'] only
                    written = ~np.isnan(ows[:-1])
                    assert np.array_equal(gws[:-1][written].view(np.uint8), ows[:-1][written].view(np.uint8)), (method, plan)
        planner.close()


def test_planner_autotune_and_plan_cache(oracle, handle):
    import torch
    import rtatblas_b200 as rt
    from gpu_util import dev

    planner = rt.Planner("gemm", "double")
    m, n, k = 300, 200, 100
    A, B = dev(oracle.fill(m * k, 1)), dev(oracle.fill(k * n, 2))
    Cm = torch.empty(m * n, dtype=torch.float64, device="cuda")
    need = max(planner.workspace(0, 0, m, n, k, p) for p in planner.plans())
    ws = torch.empty(need // 8 + 1, dtype=torch.float64, device="cuda")
    used = [planner.gemm(handle, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, Cm, m, ws, need) for _ in range(20)]
    assert used[:8] == planner.plans()          # exploration in enumeration order (planning_system.h:72-76)
    assert len(set(used[9:])) == 1              # converged
    saved = planner.save_plans()
    assert saved[0]["key"] == {"transA": "N", "transB": "N", "m": m, "n": n, "k": k}
    fresh = rt.Planner("gemm", "double")
    assert fresh.load_plans(saved) == 1
    assert fresh.gemm(handle, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, Cm, m, ws, need) == used[-1]
    # a plan that does not fit the workspace degrades to the default plan instead of failing (planning_system.h:102-105)
    stats_before = sum(len(o["times"]) for p in fresh.statistics() for o in p["options"])
    fresh.gemm(handle, 0, 0, m, n, k, 1.0, A, m, B, k, 0.0, Cm, m, None, 0, plan="TTT")
    stats = fresh.statistics()
    assert sum(len(o["times"]) for p in stats for o in p["options"]) == stats_before + 1
    assert any(o["option"] == {"transA": "N", "transB": "N", "transC": "N"} and o["times"] for o in stats[0]["options"])
