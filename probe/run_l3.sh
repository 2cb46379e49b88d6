#!/bin/bash
# SYRK / TRSM bring-up on the GPU box: parity tests, the restated reference executor tests, goldens from the
# unmodified reference, and quick timings against the reference's cuBLAS calls.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_l3.py -x -q -m gpu 2>&1 | tail -40 > gpurun_out/l3_pytest.log
(timeout 300 build/bin/gpu_tests SYRK; timeout 300 build/bin/gpu_tests TRSM; timeout 60 build/bin/gpu_tests Facade) > gpurun_out/l3_cpp.log 2>&1
timeout 300 bash tests/golden/make_golden.sh > gpurun_out/l3_golden.log 2>&1
timeout 600 python probe/perf_l3.py > gpurun_out/l3_perf.log 2>&1
tail -5 gpurun_out/l3_pytest.log; tail -8 gpurun_out/l3_cpp.log; tail -3 gpurun_out/l3_golden.log; cat gpurun_out/l3_perf.log
