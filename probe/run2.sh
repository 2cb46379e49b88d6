#!/bin/bash
mkdir -p gpurun_out
bash tests/golden/make_golden.sh > gpurun_out/golden.log 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python probe/perf_quick.py all > gpurun_out/perf_quick.log 2>&1
tail -30 gpurun_out/pytest_gpu.log; cat gpurun_out/perf_quick.log; cat gpurun_out/golden.log
