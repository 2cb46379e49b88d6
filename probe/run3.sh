#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "dgemm" > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python probe/perf_quick.py dgemm > gpurun_out/perf_quick.log 2>&1
tail -15 gpurun_out/pytest_gpu.log; cat gpurun_out/perf_quick.log
