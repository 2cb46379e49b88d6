#!/bin/bash
mkdir -p gpurun_out/edges
python probe/perf_edges.py > gpurun_out/edges/perf_edges.txt 2>&1; cat gpurun_out/edges/perf_edges.txt
python -m pytest tests -x -q -m gpu 2>&1 | tail -5 | tee gpurun_out/edges/pytest_gpu.txt
