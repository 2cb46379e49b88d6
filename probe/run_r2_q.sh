#!/bin/bash
# Round 2, call Q (2 GPUs): the column-split bench line, the NCCL-broadcast fallback transport once on hardware, the multi-GPU tests
out=gpurun_out/r2q; mkdir -p $out
nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm --format=csv > $out/gpus.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 ) > $out/bench_n2.json 2> $out/bench_n2.err; tail -c 300 $out/bench_n2.err
( time RTAT_B200_MG_TRANSPORT=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 3 --warmup 3 --no-e2e ) > $out/bench_n2_nccl.json 2> $out/bench_n2_nccl.err; tail -c 300 $out/bench_n2_nccl.err
( time timeout 900 python -m pytest tests -q -m gpu -k "multi_gpu or mg_" ) > $out/pytest_mg.txt 2>&1; tail -4 $out/pytest_mg.txt
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 3 --warmup 3 ) > $out/bench_n2_ref.json 2> $out/bench_n2_ref.err
echo done
