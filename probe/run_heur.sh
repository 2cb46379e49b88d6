#!/bin/bash
mkdir -p gpurun_out/heur
timeout 200 python probe/perf_edges.py 2>&1 | grep -A0 "auto" | tee gpurun_out/heur/auto.txt
timeout 200 python probe/perf_l3.py 2>&1 | tee gpurun_out/heur/l3.txt | grep -i trsm | head -12
RTAT_B200_E2E_OPTIONS="" timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'TF', round(d['e2e']['ms_per_step'],2), 'ms')" | tee gpurun_out/heur/e2e.txt
