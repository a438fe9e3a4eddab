import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
N, D = 4096, 128
dev = torch.device("cuda:0")
h_a = torch.empty(N * D * 3, dtype=torch.float64).pin_memory()
d_a = torch.empty(N * D * 3, dtype=torch.float64, device=dev)
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
print("H2D 12.6MB pinned ms", t(lambda: d_a.copy_(h_a, non_blocking=True)))
print("D2H 12.6MB pinned ms", t(lambda: h_a.copy_(d_a, non_blocking=True)))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d_a.copy_(h_a, non_blocking=True)
    with torch.cuda.stream(s2): h_a2.copy_(d_b, non_blocking=True)
h_a2 = torch.empty(N * D * 3, dtype=torch.float64).pin_memory(); d_b = torch.empty(N * D * 3, dtype=torch.float64, device=dev)
print("H2D || D2H 12.6MB each ms", t(both))
chunks = os.environ.get("AHMC_PIPE_CHUNKS", "8")
import ahmc_b200 as A
m, s, Minv, th, r = bench.synth(N, D, 1)
h = A.Hamiltonian(A.DiagEuclideanMetric(Minv), A.DiagGaussian(m, s))
pin = lambda a: torch.as_tensor(a).pin_memory()
thp, rp = pin(th), pin(r)
z0 = A.phasepoint(h, thp.numpy(), rp.numpy())
gp = pin(z0.lp.gradient)
z0.theta, z0.r, z0.lp.gradient = thp.numpy(), rp.numpy(), gp.numpy()
outs = [torch.empty((N, D), dtype=torch.float64).pin_memory() for _ in range(3)] + [torch.empty((N,), dtype=torch.float64).pin_memory() for _ in range(2)]
zout = A.PhasePoint(outs[0].numpy(), outs[1].numpy(), A.DualValue(outs[3].numpy(), outs[2].numpy()), A.DualValue(outs[4].numpy(), None))
plan = A.StepPlan(A.Leapfrog(0.1), h, z0, 32, out=zout)
print("chunks", chunks, "e2e call ms", t(plan, 30))
