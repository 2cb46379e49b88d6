#!/bin/bash
# Round 2, call H: whole GPU suite + smoke + the default bench line + the reference arm, then the SGEMM pair / geam tuning probes
out=gpurun_out/r2h; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $out/clocks.txt
( time timeout 2400 python -m pytest tests -x -q -m gpu ) > $out/pytest.txt 2>&1; tail -5 $out/pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > $out/smoke.txt 2>&1; tail -2 $out/smoke.txt
( time timeout 1500 python bench.py ) > $out/bench.json 2> $out/bench.err; tail -c 600 $out/bench.err
( time timeout 900 python bench.py --impl reference ) > $out/bench_ref.json 2> $out/bench_ref.err; tail -c 300 $out/bench_ref.err
timeout 120 python probe/sgemm_pair.py perf 2>&1 | tee $out/sgemm.txt
timeout 120 python probe/perf_quick.py geam_tune 2>&1 | tee $out/geam_tune.txt
echo done
