#!/bin/bash
mkdir -p gpurun_out/l2
timeout 200 python probe/perf_l2hint.py 2>&1 | tee gpurun_out/l2/times.txt
SHORT=1 timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:dgemm_ws --csv --log-file gpurun_out/l2/traffic.csv python probe/perf_l2hint.py > gpurun_out/l2/ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/l2/traffic.csv")) if len(r)>10]
hdr=rows[0]; ix={n:hdr.index(n) for n in ("ID","Metric Name","Metric Unit","Metric Value")}
for r in rows[1:]:
    print(r[ix["ID"]], r[ix["Metric Name"]], r[ix["Metric Value"]], r[ix["Metric Unit"]])
PY
timeout 300 python -m pytest tests -x -q -m gpu -k "ragged or l2_eviction or host_entry" 2>&1 | tail -3
