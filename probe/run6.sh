#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -k "sgemm" --timeout 120 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python probe/perf_quick.py sgemm 2>&1 | tail -6
