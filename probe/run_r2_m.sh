#!/bin/bash
out=gpurun_out/r2m; mkdir -p $out
for v in epi1 epi4; do echo "== $v" | tee -a $out/beta.txt; RTAT_B200_LIB=build/variants/librtat_b200_$v.so timeout 300 python probe/perf_quick.py beta 2>&1 | grep "prefetch_c=1" | tee -a $out/beta.txt; done
echo "== shipped (2)" | tee -a $out/beta.txt; timeout 300 python probe/perf_quick.py beta 2>&1 | grep "prefetch_c=1" | tee -a $out/beta.txt
timeout 300 python probe/perf_quick.py tiles 2>&1 | tee $out/tiles.txt
( time timeout 2400 python -m pytest tests -q -m gpu ) > $out/pytest.txt 2>&1; tail -8 $out/pytest.txt
( time timeout 1500 python bench.py ) > $out/bench.json 2> $out/bench.err; tail -c 600 $out/bench.err
echo done
