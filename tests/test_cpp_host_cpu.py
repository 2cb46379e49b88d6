"""Runs the host-only C++ test binary (tests/cpp/host_tests.cpp): workspace, JSON, keys, options,
option filter, plan algebra / workspace accounting, facade + plan cache — restated from the
reference's tests/workspace_test.cpp and tests/json_test.cpp.  No GPU needed."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _binary(name):
    p = os.path.join(ROOT, "build", "bin", name)
    if not os.path.exists(p):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "rtatblas_b200", "host")])
    return p


def test_host_cpp_suites():
    r = subprocess.run([_binary("host_tests")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-4000:] + r.stderr[-2000:]
    assert " 0 failures" in r.stdout
    assert r.stdout.count("[  OK  ]") >= 15


def test_host_library_exports_declared_symbols():
    import re
    import ctypes

    import rtatblas_b200 as rt

    L = ctypes.CDLL(rt.host_lib_path())
    text = open(os.path.join(ROOT, "include", "rtat_b200_host.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(rtat_host_[a-z_]+)\s*\(", text))
    assert len(names) >= 12
    assert not [n for n in names if not hasattr(L, n)]


def test_planner_workspace_matches_oracle_accounting(oracle):
    """Planning_System::calculate_workspace through the C++ op tree == oracle's closed form, all plans."""
    import rtatblas_b200 as rt
    from oracle_binding import GEMM_PAD_PLANS, GEMM_PLANS

    for method, plans, dt, elem in (("gemm", GEMM_PLANS, "double", 8), ("gemm_pad", GEMM_PAD_PLANS, "float", 4)):
        p = rt.Planner(method, dt)
        assert p.plans() == plans
        for ta in (0, 1):
            for tb in (0, 1):
                for (m, n, k) in [(70, 45, 62), (4097, 3001, 2049)]:
                    for plan in plans:
                        assert p.workspace(ta, tb, m, n, k, plan) == oracle.plan_workspace(plan, ta, tb, m, n, k, elem)
        p.close()
