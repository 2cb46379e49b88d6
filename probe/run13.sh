#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|rc=" gpurun_out/pytest_gpu.log | head -20
for w in c1 c4; do timeout 600 python bench.py --steps 20 --warmup 3 --workload $w --no-cpu-baseline 2>gpurun_out/bench_$w.err > gpurun_out/bench_$w.log; python - <<PY
import json
d=json.loads(open("gpurun_out/bench_$w.log").read().strip().splitlines()[-1])
print("$w", "value %.2f"%d["value"], "ms %.4f"%d["ms_per_step"], "plan", d["config"]["plan"], "kernel", d["roofline"]["kernel"], "roof %.3f"%d["roofline"]["frac"], "e2e %.2f (%.2f ms)"%(d["e2e"]["value"], d["e2e"]["ms_per_step"]), "launches", d["gpu_launches"], d["config"]["best_ms_per_plan"])
PY
done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
