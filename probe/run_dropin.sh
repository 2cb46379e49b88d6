#!/bin/bash
# Drop-in evidence: the reference's own host code over our C ABI (tests/test_gpu_dropin.py), the same run_tests inputs through
# run_tests_b200 vs the cuBLAS-linked run_tests, (compute-sanitizer is closed on this pool).
mkdir -p gpurun_out/dropin
python -m pytest tests/test_gpu_dropin.py -x -q 2>&1 | tail -15 > gpurun_out/dropin/pytest.txt
cat gpurun_out/dropin/pytest.txt
for c in c1 c2 c3 c4; do
  oracle/_ref/run_tests_b200 probe/inputs/$c.json > gpurun_out/dropin/run_tests_b200_$c.json 2> gpurun_out/dropin/run_tests_b200_$c.err
  oracle/_ref/run_tests probe/inputs/$c.json > gpurun_out/dropin/run_tests_cublas_$c.json 2> gpurun_out/dropin/run_tests_cublas_$c.err
done
python - <<'PY'
import json
for c in ("c1","c2","c3","c4"):
    row=[]
    for which in ("b200","cublas"):
        t=open(f"gpurun_out/dropin/run_tests_{which}_{c}.json").read()
        d=json.loads(t[t.index("{"):])
        p=d["problems"][0]
        best=min((min(o["times"]),"".join(o["option"].values())) for o in p["options"])
        row.append((which,round(best[0],4),best[1]))
    print(c,row, "ratio cublas/b200 = %.3f"%(row[1][1]/row[0][1]))
PY
