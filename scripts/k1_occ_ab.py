import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, ahmc_b200 as A, bench
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev)
ctx = A.get_context(0, stream=stream.cuda_stream)
with torch.cuda.stream(stream):
    for N in (4096, 8192, 16384):
        m, s, Minv, th, r = bench.synth(N, 128, 1)
        h = A.Hamiltonian(A.DiagEuclideanMetric(Minv), A.DiagGaussian(m, s))
        z0 = A.phasepoint(h, torch.as_tensor(th, device=dev), torch.as_tensor(r, device=dev))
        plan = A.StepPlan(A.Leapfrog(0.1), h, z0, 32, flags=A.FLAG_ASYNC)
        flush = torch.zeros(512 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
        for _ in range(5):
            flush.max(); plan()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(30)]
        for a, b in ev:
            flush.max(); a.record(stream); plan(); b.record(stream)
        torch.cuda.synchronize()
        ms = sorted(a.elapsed_time(b) for a, b in ev)
        print("occ", os.environ.get("AHMC_K1_OCC", "default"), "N", N, "us min %.2f med %.2f" % (ms[0] * 1e3, ms[15] * 1e3), flush=True)
