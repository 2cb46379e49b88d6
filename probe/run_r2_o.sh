#!/bin/bash
out=gpurun_out/r2o; mkdir -p $out
timeout 300 python probe/perf_quick.py c3pad 2>&1 | tee $out/c3.txt
timeout 300 python probe/perf_quick.py tiles 2>&1 | tee $out/tiles.txt
( time timeout 2400 python -m pytest tests -q -m gpu ) > $out/pytest.txt 2>&1; tail -8 $out/pytest.txt
( time timeout 1500 python bench.py ) > $out/bench.json 2> $out/bench.err; tail -c 300 $out/bench.err
echo done
