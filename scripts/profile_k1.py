"""Small driver for ncu: headline K1 launches (L2 cold), HBM-honest single-step launches, one K2 and one K3 launch."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ahmc_b200 as A
import bench

dev = torch.device("cuda:0")
m, s, Minv, th, r = bench.synth(4096, 128, 1)
h = A.Hamiltonian(A.DiagEuclideanMetric(Minv), A.DiagGaussian(m, s))
z0 = A.phasepoint(h, torch.as_tensor(th, device=dev), torch.as_tensor(r, device=dev))
plan = A.StepPlan(A.Leapfrog(0.1), h, z0, 32)
flush = torch.zeros(64 * 1024 * 1024, dtype=torch.float64, device=dev)
for _ in range(4):
    flush.max(); torch.cuda.synchronize(); plan()
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "honest"):
    Nh = 1 << 20
    st = torch.as_tensor(s, device=dev)
    zh = A.phasepoint(h, torch.randn((Nh, 128), dtype=torch.float64, device=dev) * st, torch.randn((Nh, 128), dtype=torch.float64, device=dev) / st)
    p1 = A.StepPlan(A.Leapfrog(0.1), h, zh, 1, out=zh)
    for _ in range(3):
        p1()
if which in ("all", "k2"):
    kern = A.HMCKernel(A.Trajectory(A.EndPointTS, A.Leapfrog(0.1), A.FixedNSteps(32)))
    rng = A.PhiloxRNG(1)
    for _ in range(3):
        A.transition(rng, h, kern, z0)
if which in ("all", "k3"):
    kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.1), A.GeneralisedNoUTurn()))
    rng = A.PhiloxRNG(1)
    for _ in range(3):
        tr = A.transition(rng, h, kern, z0)
    print("nuts depth mean", tr.stat["tree_depth"].double().mean().item(), "steps", tr.stat["n_steps"].double().mean().item())
if which in ("all", "k4"):
    rng4 = np.random.Generator(np.random.PCG64(bench.SEED))
    Q, _ = np.linalg.qr(rng4.normal(size=(128, 128)))
    lam = np.exp(np.linspace(np.log(0.1), np.log(10.0), 128))
    hd = A.Hamiltonian(A.DiagEuclideanMetric(np.diag((Q * lam) @ Q.T).copy()), A.DenseGaussian(np.zeros(128), (Q / lam) @ Q.T))
    zd = A.phasepoint(hd, torch.as_tensor(th, device=dev), torch.as_tensor(r, device=dev))
    pd = A.StepPlan(A.Leapfrog(0.02), hd, zd, 32)
    for _ in range(3):
        pd()
if which in ("all", "hostlane"):
    pin = lambda a: torch.as_tensor(a).pin_memory()
    thp, rp = pin(th), pin(r)
    zh0 = A.phasepoint(h, thp.numpy(), rp.numpy())
    gp = pin(zh0.lp.gradient)
    zh0.theta, zh0.r, zh0.lp.gradient = thp.numpy(), rp.numpy(), gp.numpy()
    outs = [torch.empty((4096, 128), dtype=torch.float64).pin_memory() for _ in range(3)] + [torch.empty((4096,), dtype=torch.float64).pin_memory() for _ in range(2)]
    zout = A.PhasePoint(outs[0].numpy(), outs[1].numpy(), A.DualValue(outs[3].numpy(), outs[2].numpy()), A.DualValue(outs[4].numpy(), None))
    ph = A.StepPlan(A.Leapfrog(0.1), h, zh0, 32, out=zout)
    for _ in range(3):
        ph()
torch.cuda.synchronize()
print("done")
