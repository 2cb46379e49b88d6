#!/bin/bash
out=gpurun_out/r2v; mkdir -p $out
( time timeout 2400 python -m pytest tests -q -m gpu ) > $out/pytest.txt 2>&1; tail -6 $out/pytest.txt
( time timeout 1500 python bench.py ) > $out/bench.json 2> $out/bench.err; tail -c 300 $out/bench.err
python probe/ncu_c4_pair.py > $out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:sgemm_tcgen05 -s 2 -c 2 -o $out/r02_sgemm -f python probe/ncu_c4_pair.py > $out/ncu_sgemm.log 2>&1
tail -2 $out/ncu_sgemm.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config > $out/plain3.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/r02_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-per-config > $out/ncu_list.log 2>&1
tail -2 $out/ncu_list.log
echo done
