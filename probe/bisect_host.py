import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    import rtatblas_b200 as rt
    mode, n, variant, sk, ta = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
    h = rt.Handle(stream=torch.cuda.current_stream())
    h.set_option("dgemm.variant", variant); h.set_option("dgemm.streamk", sk)
    if mode == "host":
        hA = torch.rand(n * n, dtype=torch.float64).pin_memory(); hB = torch.rand(n * n, dtype=torch.float64).pin_memory(); hC = torch.empty(n * n, dtype=torch.float64).pin_memory()
        h.gemm(True, ta, 0, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n, host=True)
    elif mode == "oneblock":
        big = torch.rand(3 * n * n, dtype=torch.float64, device="cuda")
        h.gemm(True, ta, 0, n, n, n, 1.0, big[: n * n], n, big[n * n : 2 * n * n], n, 0.0, big[2 * n * n :], n)
    elif mode == "cudamalloc":
        import ctypes
        rtl = ctypes.CDLL("libcudart.so.12")
        p = ctypes.c_void_p(); assert rtl.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(3 * n * n * 8)) == 0
        base = p.value
        rtl.cudaMemset(p, 0, ctypes.c_size_t(3 * n * n * 8))
        h.gemm(True, ta, 0, n, n, n, 1.0, base, n, base + n * n * 8, n, 0.0, base + 2 * n * n * 8, n)
    elif mode.startswith("seq"):
        # mimic bench.py: planner explores all plans on device buffers, then the host entry
        A = torch.rand(n * n, dtype=torch.float64, device="cuda"); B = torch.rand(n * n, dtype=torch.float64, device="cuda"); Cm = torch.empty(n * n, dtype=torch.float64, device="cuda")
        planner = rt.Planner("gemm", "double")
        plans = planner.plans() if mode == "seq" else mode[3:].split(",")
        need = max(planner.workspace(ta, 0, n, n, n, p) for p in plans)
        ws = torch.empty(max(need, 8) // 8, dtype=torch.float64, device="cuda")
        for p in plans:
            planner.gemm(h, ta, 0, n, n, n, 1.0, A, n, B, n, 0.0, Cm, n, ws, need, plan=p)
        torch.cuda.synchronize()
        print("planner phase ok", flush=True)
        exp = os.environ.get("EXP", "host")
        if exp == "host":
            hA = torch.empty(n * n, dtype=torch.float64, pin_memory=True); hB = torch.empty(n * n, dtype=torch.float64, pin_memory=True); hC = torch.empty(n * n, dtype=torch.float64, pin_memory=True)
            hA.copy_(A); hB.copy_(B)
            torch.cuda.synchronize()
            h.gemm(True, ta, 0, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n, host=True)
        elif exp == "sk_off_then_on":
            hA = torch.rand(n * n, dtype=torch.float64); hB = torch.rand(n * n, dtype=torch.float64); hC = torch.empty(n * n, dtype=torch.float64)
            h.set_option("dgemm.streamk", 1)
            h.gemm(True, ta, 0, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n, host=True)
        elif exp == "host_twice":
            hA = torch.rand(n * n, dtype=torch.float64); hB = torch.rand(n * n, dtype=torch.float64); hC = torch.empty(n * n, dtype=torch.float64)
            h.set_option("dgemm.streamk", 0)
            h.gemm(True, ta, 0, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n, host=True)
            torch.cuda.synchronize(); print("first host call ok", flush=True)
            h.set_option("dgemm.streamk", 1)
            h.gemm(True, ta, 0, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n, host=True)
        elif exp == "device_sk_after_host":
            hA = torch.rand(n * n, dtype=torch.float64); hB = torch.rand(n * n, dtype=torch.float64); hC = torch.empty(n * n, dtype=torch.float64)
            h.set_option("dgemm.streamk", 0)
            h.gemm(True, ta, 0, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n, host=True)
            torch.cuda.synchronize(); print("first host call ok", flush=True)
            h.set_option("dgemm.streamk", 1)
            h.gemm(True, ta, 0, n, n, n, 1.0, A, n, B, n, 0.0, Cm, n)
        elif exp == "pageable":
            hA = torch.rand(n * n, dtype=torch.float64); hB = torch.rand(n * n, dtype=torch.float64); hC = torch.empty(n * n, dtype=torch.float64)
            h.gemm(True, ta, 0, n, n, n, 1.0, hA, n, hB, n, 0.0, hC, n, host=True)
        elif exp == "pinned_then_device":
            hA = torch.empty(n * n, dtype=torch.float64, pin_memory=True)
            hA.copy_(A)
            torch.cuda.synchronize()
            h.gemm(True, ta, 0, n, n, n, 1.0, A, n, B, n, 0.0, Cm, n)
        elif exp == "newdevice":
            A2 = torch.rand(n * n, dtype=torch.float64, device="cuda"); B2 = torch.rand(n * n, dtype=torch.float64, device="cuda"); C2 = torch.empty(n * n, dtype=torch.float64, device="cuda")
            h.gemm(True, ta, 0, n, n, n, 1.0, A2, n, B2, n, 0.0, C2, n)
        elif exp == "cudamalloc":
            import ctypes
            rtl = ctypes.CDLL("libcudart.so.12")
            p = ctypes.c_void_p(); assert rtl.cudaMalloc(ctypes.byref(p), ctypes.c_size_t(3 * n * n * 8)) == 0
            base = p.value
            rtl.cudaMemset(p, 0, ctypes.c_size_t(3 * n * n * 8))
            h.gemm(True, ta, 0, n, n, n, 1.0, base, n, base + n * n * 8, n, 0.0, base + 2 * n * n * 8, n)
        elif exp == "direct_again":
            for _ in range(3):
                h.gemm(True, ta, 0, n, n, n, 1.0, A, n, B, n, 0.0, Cm, n)
    torch.cuda.synchronize()
    print("OK", h.last_kernel())
    sys.exit(0)
for exp in ["sk_off_then_on", "host_twice", "device_sk_after_host"]:
  for args in [("seqNNN", 1024, 0, 0, 0), ("seqNNN", 1024, 0, 1, 0)]:
    os.environ["EXP"] = exp
    print(exp, end=" ")
    r = subprocess.run([sys.executable, __file__] + [str(a) for a in args], capture_output=True, text=True)
    print(args, "rc", r.returncode, r.stdout.strip().splitlines(), "|", [l for l in r.stderr.splitlines() if "rtat" in l or "Error" in l][:2], flush=True)
