#!/usr/bin/env python
"""bench.py — headline benchmark of the rtatblas GEMM execution path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload auto|c1..c5]

One "step" is one tuned GEMM of the workload through the reference-facing API
(Planning_System::create_plan + ::execute, i.e. rtat.h's surface via include/rtat_b200_host.h) on
synthetic inputs already resident in HBM.  Rank 0 prints ONE JSON line (see README / DESIGN.md §5).

Workloads (BASELINE.json configs): N = 1 defaults to config 2 (DGEMM 8192^3 op TN, the configuration
the metric is quoted on: planner picks between the explicit-transpose plan TNN and the fused plan
NNN); N > 1 runs config 5 (DGEMM 32768^3, C's columns split across the N GPUs, A broadcast over
NCCL each step) and reports strong scaling.

--impl reference runs the UNMODIFIED reference (oracle/_ref/run_tests = /root/reference/src compiled
in place against cuBLAS) on the same problem on this GPU and reports its best plan; if that binary
is missing it falls back to the host-CPU BLAS baseline.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (dtype, transA, transB, m, n, k)
    "c1": ("double", "N", "N", 1024, 1024, 1024),
    "c2": ("double", "T", "N", 8192, 8192, 8192),
    "c3": ("double", "N", "T", 4097, 3001, 2049),
    "c4": ("float", "N", "N", 131072, 128, 4096),
    "c5": ("double", "N", "N", 32768, 32768, 32768),
}
DMMA_PEAK_TFLOPS = 37.0  # measured FP64 DMMA issue peak on this pool's B200 (profiles/r01_probe_fp64_rates.txt)


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        # only the samples taken under load (upper half) describe the timed region
        sm_load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": statistics.median(sm_load) if sm_load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_blas_baseline(cfg, budget_s=12.0):
    """Host BLAS (numpy -> OpenBLAS) on a bounded k-slab of the same problem, all host cores."""
    import numpy as np

    dtype, ta, tb, m, n, k = cfg
    npdt = np.float64 if dtype == "double" else np.float32
    try:
        from threadpoolctl import threadpool_info

        cores = max([p.get("num_threads", 1) for p in threadpool_info()] or [os.cpu_count()])
    except Exception:
        cores = os.cpu_count()
    ms, ns = min(m, 8192), min(n, 8192)
    ks = min(k, 512)
    rng = np.random.default_rng(0)
    A = rng.random((ms, ks), dtype=npdt) if ta == "N" else rng.random((ks, ms), dtype=npdt).T
    B = rng.random((ks, ns), dtype=npdt) if tb == "N" else rng.random((ns, ks), dtype=npdt).T
    A @ B  # warm-up (thread pool spin-up)
    t0, reps = time.perf_counter(), 0
    while True:
        A @ B
        reps += 1
        if time.perf_counter() - t0 > budget_s or reps >= 5:
            break
    dt = (time.perf_counter() - t0) / reps
    return {"value": 2.0 * ms * ns * ks / dt * 1e-12, "unit": "TFLOP/s", "cores": cores, "kind": "port",
            "sample": f"numpy/OpenBLAS {dtype} GEMM op {ta}{tb} on a {ms}x{ns}x{ks} slab of the {m}x{n}x{k} problem, {reps} reps, all host threads"}


def cpu_oracle_baseline(cfg, budget_s=12.0):
    """The reference's only CPU implementation of this path — the element-wise triple loop its tests check the GPU
    against (tests/common.h:158-193), restated in oracle/oracle.c (OpenMP over columns) — timed on a bounded block of
    the same problem: a square block of C with the FULL k extent.  The host BLAS figure rides along as `host_blas`."""
    import numpy as np

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_binding

    orc = oracle_binding.load()
    dtype, ta, tb, m, n, k = cfg
    npdt = np.float64 if dtype == "double" else np.float32
    cores = os.cpu_count() or 1
    # size the block for roughly budget_s / 2 at ~0.25 GFLOP/s per core
    side = int(max(64, min(m, n, (0.5 * budget_s * 0.25e9 * cores / (2.0 * k)) ** 0.5)) // 64 * 64) or min(m, n)
    ms, ns = min(m, side), min(n, side)
    ar, ac = (k, ms) if ta == "T" else (ms, k)
    br, bc = (ns, k) if tb == "T" else (k, ns)
    A, B = orc.fill(ar * ac, 1, npdt), orc.fill(br * bc, 2, npdt)
    Cm = np.zeros(ms * ns, dtype=npdt)
    t0, reps = time.perf_counter(), 0
    while True:
        orc.gemm(int(ta == "T"), int(tb == "T"), ms, ns, k, 1.0, A, ar, B, br, 0.0, Cm, ms)
        reps += 1
        if time.perf_counter() - t0 > budget_s / 2 or reps >= 3:
            break
    dt = (time.perf_counter() - t0) / reps
    out = {"value": 2.0 * ms * ns * k / dt * 1e-12, "unit": "TFLOP/s", "cores": cores, "kind": "port",
           "sample": f"oracle_{'d' if dtype == 'double' else 's'}gemm (the reference's test_gemm triple loop, OpenMP) on the {ms}x{ns} leading block of C "
                     f"of the {m}x{n}x{k} op {ta}{tb} problem, full k, {reps} reps"}
    out["host_blas"] = {k2: v for k2, v in cpu_blas_baseline(cfg, budget_s=budget_s / 2).items() if k2 in ("value", "unit", "cores", "sample")}
    return out


# ---------------------------------------------------------------------------------------------------
def run_reference(args, cfg_name, cfg, rank, world):
    """The reference's own driver on this GPU (cuBLAS-backed executor), best plan."""
    dtype, ta, tb, m, n, k = cfg
    if rank != 0:
        return
    sample_note = ""
    if world > 1:
        # the reference is a single-GPU program: give it one rank's column block of the split problem
        n = n // world
        sample_note = f" (one rank's column block, n/{world})"
    flops = 2.0 * m * n * k
    reps = args.steps + args.warmup if flops < 5e12 else min(args.steps + args.warmup, 3)
    binary = os.path.join(ROOT, "oracle", "_ref", "run_tests")
    line = {"impl": "reference", "metric": "gemm_tflops", "unit": "TFLOP/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup,
            "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64" if dtype == "double" else "f32",
            "data": "synthetic", "config": {"workload": f"{cfg_name}: {dtype} GEMM op {ta}{tb} m={m} n={n} k={k}" + sample_note + " through the reference's run_tests (exhaustive, all plans)"}}
    if os.path.exists(binary) and (m * k + k * n + m * n) * (8 if dtype == "double" else 4) < 60e9:
        doc = {"keywords": {"run_type": "exhaustive", "method": "gemm", "data_type": dtype, "repetitions": reps},
               "problems": [{"transA": ta, "transB": tb, "m": m, "n": n, "k": k}]}
        with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
            json.dump(doc, f)
        try:
            out = subprocess.run([binary, f.name], capture_output=True, text=True, timeout=1500)
            res = json.loads(out.stdout)
            best = None
            for o in res["problems"][0]["options"]:
                ts = o["times"][min(args.warmup, reps - 1):] or o["times"]
                mean = sum(ts) / len(ts)
                name = "".join(o["option"][x] for x in ("transA", "transB", "transC"))
                if best is None or mean < best[0]:
                    best = (mean, name)
            v = flops / (best[0] * 1e-3) * 1e-12
            line.update({"value": v, "ms_per_step": best[0],
                         "cpu_baseline": {"value": v, "unit": "TFLOP/s", "cores": 0, "kind": "reference",
                                          "sample": f"unmodified reference executor (oracle/_ref/run_tests, cuBLAS 12.9) on this GPU, best plan {best[1]}; the reference has no CPU implementation of the path"},
                         "e2e": {"value": v, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
            line["config"]["world_note"] = "single-GPU program: the reference has no multi-GPU path" if world > 1 else ""
            print(json.dumps(line), flush=True)
            return
        except Exception as e:  # fall through to the CPU baseline
            line["config"]["reference_gpu_run_failed"] = str(e)[:200]
        finally:
            os.unlink(f.name)
    cb = cpu_blas_baseline(cfg, budget_s=20.0)
    line.update({"value": cb["value"], "ms_per_step": flops / (cb["value"] * 1e12) * 1e3, "cpu_baseline": cb,
                 "e2e": {"value": cb["value"], "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto"] + sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cfg_name = args.workload if args.workload != "auto" else ("c2" if world == 1 else "c5")
    cfg = CONFIGS[cfg_name]

    if args.impl == "reference":
        run_reference(args, cfg_name, cfg, rank, world)
        return

    import torch

    import rtatblas_b200 as rt

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no GPU visible; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    dtype, ta_s, tb_s, m, n, k = cfg
    ta, tb = int(ta_s == "T"), int(tb_s == "T")
    dbl = dtype == "double"
    tdt = torch.float64 if dbl else torch.float32
    esz = 8 if dbl else 4
    # column split of C (and of op(B)) across ranks: rank r owns columns [r*n/world, (r+1)*n/world)
    n_loc = n // world
    assert n_loc * world == n, "n must divide evenly across ranks"
    if world > 1:
        assert not tb, "column split needs op(B) = B so that a column block is a contiguous view"

    stream = torch.cuda.current_stream()
    handle = rt.Handle(stream=stream)
    method = "gemm_pad" if cfg_name == "c3" else "gemm"  # config 3 is the one where pad/copy plans matter: 64-plan space
    planner = rt.Planner(method, dtype)
    # three samples per plan before converging (the reference's default of one sample is easily fooled on
    # 70 us problems); an extension of Planning_System, see set_tests_until_converge
    SAMPLES = 3 if method == "gemm" else 1
    planner.set_tests_until_converge(SAMPLES)
    ar, ac = (k, m) if ta else (m, k)
    br, bc = (n_loc, k) if tb else (k, n_loc)
    B = torch.empty(br * bc, dtype=tdt, device="cuda")
    Cm = torch.empty(m * n_loc, dtype=tdt, device="cuda")
    handle.fill_uniform(B, 0xB200 + 2 + rank, -1.0, 1.0)
    plans = planner.plans()
    transport, cs, A_src, A_stage = "none", None, None, None
    if world == 1:
        A = torch.empty(ar * ac, dtype=tdt, device="cuda")
        handle.fill_uniform(A, 0xB200 + 1, -1.0, 1.0)
    else:
        assert dbl, "the column-split path is DGEMM (config 5)"

        class _DevArray:  # zero-copy torch view of memory the C library allocated
            def __init__(self, ptr, count):
                self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}

        cs = rt.ColumnSplit(handle, rank, world)
        hb = torch.zeros(rt.mg.HANDLE_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            a_ptr, hbytes = cs.alloc_shared(ar * ac * esz)  # IPC-exportable: the other ranks' copy engines pull from it
            A = torch.as_tensor(_DevArray(a_ptr, ar * ac), device="cuda")
            handle.fill_uniform(A, 0xB200 + 1, -1.0, 1.0)
            hb.copy_(torch.tensor(list(hbytes), dtype=torch.uint8))
            A_src = A
        torch.cuda.synchronize()
        dist.broadcast(hb, src=0)
        ok = torch.ones(1, device="cuda")
        if rank != 0:
            A = torch.empty(ar * ac, dtype=tdt, device="cuda")  # local replica, filled panel by panel every step
            try:
                peer = cs.open_shared(bytes(hb.cpu().numpy().tobytes()))
                A_src, A_stage = peer, A
            except rt.RtatError:
                ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        transport = "nvlink-p2p-copy-engine-pull" if ok.item() > 0 else "nccl-broadcast"
    free_b, _ = torch.cuda.mem_get_info()
    # plans whose scratch fits comfortably; the others are still explored but degrade to the default plan
    need = 0 if world > 1 else max([w for w in (planner.workspace(ta, tb, m, n_loc, k, p) for p in plans) if w < 0.6 * free_b] + [0])
    ws = torch.empty(max(need, 8) // 8, dtype=torch.float64, device="cuda")
    PANELS = 0  # automatic schedule: 1/16, 3/16, 4/16, 8/16 of K

    def step():
        if world == 1:
            return planner.gemm(handle, ta, tb, m, n_loc, k, 1.0, A, ar, B, br, 0.0, Cm, m, ws, need)
        if transport == "nccl-broadcast":
            dist.broadcast(A, src=0)  # A replicated once per step over NCCL / NVLink
            handle.gemm(dbl, ta, tb, m, n_loc, k, 1.0, A, ar, B, br, 0.0, Cm, m)
        else:
            # K-panel pipeline: copy engines pull panel p+1 of A from rank 0 while the DMMA kernel multiplies panel p
            cs.dgemm(ta, tb, m, n_loc, k, 1.0, A_src, ar, A_stage, B, br, 0.0, Cm, m, panels=PANELS)
        return "NNN"

    # ---- tuning (untimed): the planner explores every plan once, then converges ------------------------
    tune = [step() for _ in range(SAMPLES * len(plans) + 1 if world == 1 else 1)]
    torch.cuda.synchronize()
    chosen = tune[-1]
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = handle.launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = handle.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    kernel_name = handle.last_kernel()
    if dist:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps
    flops = 2.0 * m * n * k
    value = flops / (ms_per_step * 1e-3) * 1e-12

    # ---- roofline: the GEMM kernel alone, CUDA events on its stream ------------------------------------
    if world > 1 and rank != 0 and transport != "nccl-broadcast":
        torch.cuda.synchronize()  # A (the local replica) is complete after the last step
    k_ms = []
    for _ in range(min(args.steps, 5)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        handle.gemm(dbl, ta, tb, m, n_loc, k, 1.0, A, ar, B, br, 0.0, Cm, m)
        b.record()
        b.synchronize()
        k_ms.append(a.elapsed_time(b))
    k_avg = sum(k_ms) / len(k_ms)
    achieved = 2.0 * m * n_loc * k / (k_avg * 1e-3) * 1e-12
    peaks = measured_peaks()
    if dbl:
        # DRAM bytes per launch of this kernel from the committed ncu --set full capture (profiles/r01_dgemm_ws_ncu_summary.csv)
        traffic = {"c2": 9.668807e9 + 544.913664e6, "c1": 0.016827e9 + 0.002816e6, "c3": 0.385148e9 + 97.243648e6}.get(cfg_name)
        roof = {"bound": "tensor", "achieved": achieved, "peak": DMMA_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / DMMA_PEAK_TFLOPS,
                "traffic": traffic, "algorithmic_bytes": 8.0 * (m * k + k * n_loc + m * n_loc), "kernel": handle.last_kernel(),
                "peak_source": "FP64 DMMA.8x8x4 issue peak measured on this pool (profiles/r01_probe_fp64_rates.txt); MEASURED_PEAKS.json has no FP64 entry",
                "hbm_gbs_measured": peaks.get("hbm_gbs")}
    else:
        tf32 = peaks.get("bf16_tflops", 1590.0) / 2.0 / 3.0  # 3xTF32: a third of the TF32 rate, itself half of bf16
        roof = {"bound": "tensor", "achieved": achieved, "peak": tf32, "unit": "TFLOP/s", "frac": achieved / tf32, "traffic": None,
                "kernel": handle.last_kernel(), "peak_source": "MEASURED_PEAKS.json bf16_tflops / 2 (TF32) / 3 (3xTF32 split products)"}

    # ---- end to end: host buffers through the C ABI, H2D + GEMM + D2H inside the timed region ------------
    e2e = None
    if not args.no_e2e:
        # Every rank is on this node, so A (one copy per process here, as a stand-in for a shared host mapping) is read
        # by every GPU over its own PCIe link: each rank runs the host entry point on its column block, which
        # pipelines K-panels of A, its block of B and its block of C against the GEMM.
        hA = torch.empty(ar * ac, dtype=tdt, pin_memory=True)
        hB = torch.empty(br * bc, dtype=tdt, pin_memory=True)
        hC = torch.empty(m * n_loc, dtype=tdt, pin_memory=True)
        hA.copy_(A)  # N > 1: the local replica is complete after the timed steps above
        hB.copy_(B)
        torch.cuda.synchronize()
        e_steps = max(1, min(args.steps, 3))

        api = "rtat_b200_dgemm_host (pinned host buffers)"
        h2d = (ar * ac + br * bc) * esz
        # experiments: "percall" = every rank uploads all of A; "push" = flags pushed to the peers (stream memops wait); "spin" = push + polling kernels
        mg_mode = os.environ.get("RTAT_B200_MG_E2E", "")
        if mg_mode in ("push", "spin"):
            handle.set_option("mg.signal", 0)
        if mg_mode == "spin":
            handle.set_option("mg.spin_wait", 1)
        if os.environ.get("RTAT_B200_HOST_TRACE"):  # per-stage timeline of every host-entry call on stderr
            handle.set_option("host.trace", 1)
        for kv in filter(None, os.environ.get("RTAT_B200_E2E_OPTIONS", "").split(",")):  # experiments: "host.lead=3,dgemm.tile=64"
            key, val = kv.split("=")
            handle.set_option(key.strip(), int(val))
        if world > 1 and transport != "nccl-broadcast" and mg_mode != "percall":
            # host-buffer column split: every rank uploads 1/world of each K-panel of A into its replica and pulls the other
            # shares from the peers' replicas over NVLink (copy engines, flag-synchronised) — A crosses PCIe once per step
            # release the device-resident A of the HBM-resident bench above: the host entry owns its own replicas
            torch.cuda.synchronize()
            if rank != 0:
                cs.close_shared(A_src)
            A = A_src = A_stage = None
            dist.barrier()
            if rank == 0:
                cs.free_shared(a_ptr)
            torch.cuda.empty_cache()
            hbytes = cs.host_setup(ar * ac * esz)
            mine = torch.tensor(list(hbytes), dtype=torch.uint8, device="cuda")
            allh = torch.empty(rt.mg.HANDLE_BYTES * world, dtype=torch.uint8, device="cuda")
            dist.all_gather_into_tensor(allh, mine)
            cs.host_connect(allh.cpu().numpy().tobytes())
            api = ("rtat_b200_mg_dgemm_host (pinned host buffers): per K-panel every rank uploads its 1/N share of A and pulls the "
                   "others over NVLink; B/C column blocks over the rank's own PCIe link")
            h2d = (ar * ac + br * bc * world) * esz

            def e2e_step():
                cs.dgemm_host(ta, tb, m, n_loc, k, 1.0, hA, ar, hB, br, 0.0, hC, m)
        else:
            h2d = (ar * ac + br * bc) * world * esz

            def e2e_step():
                handle.gemm(dbl, ta, tb, m, n_loc, k, 1.0, hA, ar, hB, br, 0.0, hC, m, host=True)

        e2e_step()
        torch.cuda.synchronize()
        if dist:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            e2e_step()
        torch.cuda.synchronize()
        t1 = (time.perf_counter() - t0) * 1e3 / e_steps
        if dist:
            t = torch.tensor([t1], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t1 = float(t.item())
        e2e = {"value": flops / (t1 * 1e-3) * 1e-12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": m * n * esz,
               "ms_per_step": t1, "steps": e_steps, "api": api}

    if rank == 0:
        stats = planner.statistics() if world == 1 else [{"options": []}]
        per_plan = {}
        for o in stats[0]["options"]:
            name = "".join(o["option"][x] for x in ("transA", "transB", "transC", "padA", "padB", "padC") if x in o["option"])
            if o["times"]:
                per_plan[name] = round(sorted(o["times"])[len(o["times"]) // 2], 4)  # median of the logged samples
        if len(per_plan) > 8:  # 64-plan space: keep the line readable
            per_plan = dict(sorted(per_plan.items(), key=lambda kv: kv[1])[:8])
        line = {
            "metric": "gemm_tflops", "value": value, "unit": "TFLOP/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
            "dtype": "f64" if dbl else "f32", "data": "synthetic",
            "config": {"workload": f"{cfg_name}: {dtype} GEMM op {ta_s}{tb_s} m={m} n={n} k={k} alpha=1 beta=0 through Planning_System (autotuned plan)",
                       "plan": chosen, "plan_meaning": "NNN = transpose/pad fused into the GEMM's loads; T.. = explicit MatrixMove pass first",
                       "best_ms_per_plan": per_plan,
                       "parallelism": f"column split of C over {world} GPU(s)" + (f", A replicated every step from rank 0 by {transport}, 4 K-panels (1/16, 3/16, 4/16, 8/16 of K) pipelined with the GEMM" if world > 1 else ""),
                       "l2_policy": "inputs larger than L2 (126 MB): no flush needed" if (ar * ac + br * bc) * esz > 200e6 else "inputs fit in L2; back-to-back steps measure the warm-L2 case the reference's harness also sees",
                       "pct_of_fp64_dmma_peak": round(100 * value / (DMMA_PEAK_TFLOPS * world), 2) if dbl else None},
            "roofline": roof, "gpu_launches": launches, "clocks": clocks, "kernel": kernel_name,
        }
        if e2e:
            line["e2e"] = e2e
        if not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_oracle_baseline(cfg)
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
