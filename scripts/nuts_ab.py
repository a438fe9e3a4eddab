"""Time the NUTS transition kernel (4096 x 128 diag Gaussian, Philox) with CUDA events; one lib per process."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ahmc_b200 as A
import bench
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev)
A.get_context(0, stream=stream.cuda_stream)
m, s, Minv, th, r = bench.synth(4096, 128, 1)
h = A.Hamiltonian(A.DiagEuclideanMetric(Minv), A.DiagGaussian(m, s))
with torch.cuda.stream(stream):
    z0 = A.phasepoint(h, torch.as_tensor(th * s, device=dev), torch.as_tensor(r, device=dev))
    for eps in (0.1, 0.4):
        kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(eps), A.GeneralisedNoUTurn()))
        rng = A.PhiloxRNG(1)
        z = z0
        for _ in range(5):
            z = A.transition(rng, h, kern, z, flags=A.FLAG_ASYNC).z
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps, steps = 20, 0
        torch.cuda.synchronize()
        stats = []
        e0.record(stream)
        for _ in range(reps):
            tr = A.transition(rng, h, kern, z, flags=A.FLAG_ASYNC)
            z = tr.z
            stats.append(tr.stat["n_steps"])
        e1.record(stream)
        torch.cuda.synchronize()
        steps = sum(int(x.sum().item()) for x in stats)
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps(dict(lib=os.environ.get("AHMC_B200_LIB", "default"), eps=eps, ms_per_transition=ms,
                              mean_steps=steps / reps / 4096, rate=steps * 128 / (e0.elapsed_time(e1) * 1e-3))))
    # persistent launch: T transitions per launch, chains free-running
    for eps in (0.1, 0.4):
        kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(eps), A.GeneralisedNoUTurn()))
        rng = A.PhiloxRNG(1)
        T = 20
        zl, _, st = A.sample_transitions(rng, h, kern, z0, T, keep_draws=False, flags=A.FLAG_ASYNC)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        zl, _, st = A.sample_transitions(rng, h, kern, zl, T, keep_draws=False, flags=A.FLAG_ASYNC)
        e1.record(stream)
        torch.cuda.synchronize()
        steps = int(st["n_steps"].sum().item())
        print(json.dumps(dict(mode="persistent", eps=eps, ms_per_transition=e0.elapsed_time(e1) / T, mean_steps=steps / T / 4096,
                              rate=steps * 128 / (e0.elapsed_time(e1) * 1e-3))))
