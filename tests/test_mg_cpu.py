"""CPU tests of the multi-GPU column-split host logic (include/rtat_b200_mg.h): the partition and
K-panel schedules, and — with a world_size-2 gloo process group — that the ranks' blocks tile C exactly
and that the 64-byte handle exchange plumbing the GPU path relies on works."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_partition_tiles_the_columns_exactly():
    import rtatblas_b200 as rt

    for n in (0, 1, 7, 128, 3001, 32768):
        for world in (1, 2, 3, 4, 8):
            blocks = [rt.mg.partition(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f0 + c0 == f1
            assert blocks[-1][0] + blocks[-1][1] == n
            counts = [c for _, c in blocks]
            assert max(counts) - min(counts) <= 1
    assert rt.mg.partition(32768, 8, 3) == (3 * 4096, 4096)  # config 5 at 8 GPUs
    with pytest.raises(rt.RtatError):
        rt.mg.partition(10, 4, 4)


def test_panels_cover_k_in_order_and_respect_the_ktile():
    import rtatblas_b200 as rt

    for k in (1, 15, 16, 17, 2049, 4096, 32768):
        for panels in (1, 2, 7, 16, 32):
            spans = [rt.mg.panel(k, panels, p) for p in range(panels)]
            pos = 0
            for f, c in spans:
                assert f == min(pos, k) and c >= 0
                if c and f + c < k:
                    assert c % 16 == 0  # only the last non-empty panel may be ragged
                pos = f + c
            assert pos == k
    assert rt.mg.panel(32768, 16, 5) == (5 * 2048, 2048)
    # automatic schedule: 1/16, 3/16, 4/16, 8/16 of K, contiguous, k-tile aligned cuts
    assert [rt.mg.auto_panel(32768, p) for p in range(4)] == [(0, 2048), (2048, 6144), (8192, 8192), (16384, 16384)]
    for k in (1, 17, 100, 2049, 5000):
        spans = [rt.mg.auto_panel(k, p) for p in range(4)]
        assert spans[0][0] == 0 and spans[-1][0] + spans[-1][1] == k
        for (f0, c0), (f1, _) in zip(spans, spans[1:]):
            assert f0 + c0 == f1 and c0 >= 0 and (f1 % 16 == 0 or f1 == k)


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1])
import numpy as np
import torch, torch.distributed as dist
import rtatblas_b200 as rt
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
m, n, k = 37, 21, 19
# every rank regenerates the same A and the full B from seeds, keeps only its column block of B
rng = np.random.default_rng(0)
A = rng.standard_normal((m, k)); B = rng.standard_normal((k, n))
first, count = rt.mg.partition(n, world, rank)
C_loc = A @ B[:, first:first + count]                      # stands in for the per-rank GEMM
# the 64-byte IPC handle travels as a byte tensor from rank 0
h = torch.zeros(64, dtype=torch.uint8)
if rank == 0:
    h[:] = torch.arange(64, dtype=torch.uint8)
dist.broadcast(h, src=0)
assert bytes(h.numpy()) == bytes(range(64))
# gather the blocks: together they must be exactly A @ B
blocks = [None] * world
dist.all_gather_object(blocks, (first, count, C_loc))
if rank == 0:
    C = np.zeros((m, n)); covered = np.zeros(n, bool)
    for f, c, blk in blocks:
        assert not covered[f:f + c].any()
        covered[f:f + c] = True
        C[:, f:f + c] = blk
    assert covered.all() and np.allclose(C, A @ B)
    # max-over-ranks timing reduction used by bench.py
t = torch.tensor([float(rank + 1)])
dist.all_reduce(t, op=dist.ReduceOp.MAX)
assert t.item() == world
dist.destroy_process_group()
print("ok", rank)
'''


def test_column_split_flow_with_two_gloo_ranks(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = [subprocess.Popen([sys.executable, str(script), ROOT], env=dict(env, RANK=str(r)), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, f"rank {r} failed:\n{o[-3000:]}"
        assert f"ok {r}" in o
