#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -k "dgemm or mg_panel" --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAILED|Error|rc=" gpurun_out/pytest_gpu.log | head -20
timeout 300 python probe/perf_quick.py tiles 2>&1 | tail -18
