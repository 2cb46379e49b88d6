#!/bin/bash
out=gpurun_out/r2s; mkdir -p $out
( time timeout 2400 python -m pytest tests -q -m gpu ) > $out/pytest.txt 2>&1; tail -6 $out/pytest.txt
( time timeout 1500 python bench.py ) > $out/bench.json 2> $out/bench.err; tail -c 300 $out/bench.err
( time timeout 600 python bench.py --workload c2 --no-per-config --no-cpu-baseline ) > $out/bench_c2.json 2> $out/bench_c2.err
RTAT_B200_HOST_TRACE=1 timeout 600 python bench.py --workload c2 --no-per-config --no-cpu-baseline --steps 2 > $out/bench_c2_trace.json 2> $out/trace_c2.err
echo done
