"""Worker for the multi-GPU host-buffer column split (rtat_b200_mg_dgemm_host): run under torch.distributed.run with one
rank per GPU, or directly (world 1).  Every rank holds the full host A and its column block of B; the result is checked
against torch's float64 matmul on the device within the DGEMM bound 2 k eps (|A||B|), for two different A in a row (the second
call reuses the replicas, so it exercises the 'peers are done reading' flags) and for the three flag transports (mg.signal / mg.spin_wait)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import rtatblas_b200 as rt  # noqa: E402


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    handle = rt.Handle(stream=torch.cuda.current_stream())
    cs = rt.ColumnSplit(handle, rank, world)
    cases = [(0, 0, 4096, 1024 * world, 4096), (1, 0, 1000, 700 * world + 3, 3000), (0, 1, 2500, 512 * world, 2304), (0, 0, 300, 40 * world, 77)]
    max_bytes = max((k if ta else m) * (m if ta else k) * 8 for ta, _, m, _, k in cases)
    eps = np.finfo(np.float64).eps
    for signal, spin in ((1, 0), (0, 0), (0, 1)):  # flags polled in the writer's page; pushed + stream memops; pushed + polling kernels
        # the flag transport is a property of a set-up exchange (who writes which page): a fresh setup per transport
        handle.set_option("mg.signal", signal)
        handle.set_option("mg.spin_wait", spin)
        hb = cs.host_setup(max_bytes)
        if world > 1:
            mine = torch.tensor(list(hb), dtype=torch.uint8, device="cuda")
            allh = torch.empty(64 * world, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(allh, mine)
            cs.host_connect(allh.cpu().numpy().tobytes())
        else:
            cs.host_connect(None)
        for ci, (ta, tb, m, n, k) in enumerate(cases):
            first, n_loc = rt.mg.partition(n, world, rank)
            ar, ac = (k, m) if ta else (m, k)
            for rep in range(2):
                g = torch.Generator().manual_seed(1000 * ci + rep)  # the same A on every rank
                hA = torch.rand(ar * ac, dtype=torch.float64, generator=g).mul_(2).sub_(1).pin_memory()
                gB = torch.Generator().manual_seed(77 + 1000 * ci + rep)
                Bfull = torch.rand(n * k, dtype=torch.float64, generator=gB).mul_(2).sub_(1)
                if tb:  # B stored n x k: this rank's columns of op(B) are rows [first, first + n_loc) -> pack them, ldb = n_loc
                    hB = Bfull.view(k, n)[:, first:first + n_loc].contiguous().view(-1).pin_memory()
                    ldb = n_loc
                else:  # B stored k x n: a contiguous column block
                    hB = Bfull.view(n, k)[first:first + n_loc].contiguous().view(-1).pin_memory()
                    ldb = k
                ldc = m + 5
                hC = torch.full((ldc * n_loc,), 3.0, dtype=torch.float64).pin_memory()
                beta = 0.0 if rep == 0 else -0.5
                cs.dgemm_host(ta, tb, m, n_loc, k, 1.25, hA, ar, hB, ldb, beta, hC, ldc)
                # reference on the device: column-major buffers are the transposes of torch's row-major views
                dA, dB = hA.cuda(), hB.cuda()
                opA = dA.view(ac, ar) if not ta else dA.view(ac, ar).t()           # (k x m): op(A)^T
                opB = dB.view(n_loc, ldb) if not tb else dB.view(k, ldb).t()       # (n_loc x k): op(B)^T
                ref_t = opB @ opA  # (n_loc x m) = C^T
                bound = (opB.abs() @ opA.abs()) * 1.25
                want = 1.25 * ref_t + beta * 3.0
                got = hC.view(n_loc, ldc)
                err = (got[:, :m].cuda() - want).abs()
                tol = 2 * k * eps * bound + 4 * eps * want.abs() + 1e-300
                assert bool((err <= tol).all()), f"rank {rank} case {ci} rep {rep} spin {spin}: max err {err.max().item():.3e}"
                assert bool((got[:, m:] == 3.0).all()), "ldc padding rows must be untouched"
                if dist:
                    dist.barrier()
        if dist:
            dist.barrier()  # nobody is still pulling from a replica that is about to be freed
        cs.host_teardown()
    cs.close()
    handle.close()
    print(f"mg_host ok rank {rank}/{world}", flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
