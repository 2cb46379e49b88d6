#!/bin/bash
# Round-end evidence run on one B200: GPU tests, smoke, bench lines for every single-GPU config, the reference arm,
# SYRK / TRSM against cuBLAS and against the unmodified reference's run_tests.
mkdir -p gpurun_out/final
python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/final/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final/smoke.txt 2>&1
for w in c1 c2 c3 c4; do python bench.py --workload $w --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final/bench_$w.json; done
python bench.py --impl reference --steps 3 --warmup 3 2>/dev/null | tail -1 > gpurun_out/final/bench_reference_c2.json
python probe/perf_l3.py > gpurun_out/final/syrk_trsm_vs_cublas.txt 2>&1
bash probe/run_l3_ref.sh > gpurun_out/final/run_tests_syrk_trsm.txt 2>&1
python probe/perf_quick.py > gpurun_out/final/perf_quick.txt 2>&1
cat gpurun_out/final/pytest_gpu.txt gpurun_out/final/smoke.txt
python - <<'PY'
import json
for w in ("c1","c2","c3","c4","reference_c2"):
    d=json.loads(open(f"gpurun_out/final/bench_{w}.json").read())
    print(w, round(d["value"],2), "ms", round(d["ms_per_step"],4), d["config"].get("plan"), "roof", d.get("roofline",{}).get("frac"), "e2e", d.get("e2e",{}).get("value"))
PY
