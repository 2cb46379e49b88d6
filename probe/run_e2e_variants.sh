#!/bin/bash
# end-to-end (host-buffer entry) of the default bench under schedule / tile overrides
mkdir -p gpurun_out/e2e
for v in "" "dgemm.tile=64" "dgemm.tile=128" "host.lead=3" "host.lead=3,dgemm.tile=64" "host.panels=4" "host.lead=5"; do
  RTAT_B200_E2E_OPTIONS="$v" timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read()); print('[$v]', 'value', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'TF', round(d['e2e']['ms_per_step'],2), 'ms')"
done | tee gpurun_out/e2e/variants.txt
