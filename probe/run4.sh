#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.log; tail -5 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cat gpurun_out/bench_ref.log; tail -5 gpurun_out/bench_ref.err
timeout 300 python probe/perf_quick.py dgemm 2>&1 | grep -E "C1|C3" 
