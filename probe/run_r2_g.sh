#!/bin/bash
out=gpurun_out/r2g; mkdir -p $out
timeout 120 python probe/sgemm_pair.py check 2>&1 | tee $out/sgemm.txt
timeout 120 python probe/sgemm_pair.py perf 2>&1 | tee -a $out/sgemm.txt
timeout 120 python probe/perf_quick.py geam_tune 2>&1 | tee $out/geam_tune.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sgemm or geam or copy or move" 2>&1 | tail -5 | tee $out/pytest.txt
echo done
