#!/bin/bash
# the unmodified reference's run_tests (cuBLAS) and ours on the same SYRK / TRSM problem files
mkdir -p gpurun_out/l3ref
for f in syrk_d syrk_s trsm_d trsm_s; do
  timeout 600 oracle/_ref/run_tests probe/inputs/$f.json > gpurun_out/l3ref/reference_$f.json 2> gpurun_out/l3ref/reference_$f.err
  timeout 600 build/bin/run_tests probe/inputs/$f.json > gpurun_out/l3ref/ours_$f.json 2> gpurun_out/l3ref/ours_$f.err
done
python - <<'PY'
import json, glob
for f in ("syrk_d", "syrk_s", "trsm_d", "trsm_s"):
    try:
        r = json.load(open(f"gpurun_out/l3ref/reference_{f}.json")); o = json.load(open(f"gpurun_out/l3ref/ours_{f}.json"))
    except Exception as e:
        print(f, "failed:", e); continue
    for pr, po in zip(r["problems"], o["problems"]):
        best = lambda p: min((min(x["times"]), "".join(x["option"][k] for k in sorted(x["option"]))) for x in p["options"] if x["times"])
        br, bo = best(pr), best(po)
        print(f, pr["key"], f"reference best {br[0]:.3f} ms ({br[1]})  ours best {bo[0]:.3f} ms ({bo[1]})  ratio {br[0]/bo[0]:.2f}")
PY
