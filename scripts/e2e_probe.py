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
zd = A.step(A.Leapfrog(0.1), h, A.phasepoint(h, torch.as_tensor(th, device=dev), torch.as_tensor(r, device=dev)), 32)
def check(tag):
    ok = all(np.array_equal(a, b.cpu().numpy()) for a, b in [(zout.theta, zd.theta), (zout.r, zd.r), (zout.lp.value, zd.lp.value),
                                                             (zout.lk.value, zd.lk.value), (zout.lp.gradient, zd.lp.gradient)])
    print(tag, "bit-identical to the device call:", ok, flush=True)
for up, down, chunks in (("ce1", "ce", "2"), ("ce1", "direct", "2"), ("direct", "direct", "1"), ("direct", "direct", "16")):
    os.environ["AHMC_PIPE_UP"], os.environ["AHMC_PIPE_DOWN"], os.environ["AHMC_PIPE_CHUNKS"] = up, down, chunks
    print("up", up, "down", down, "chunks", chunks, "e2e call ms %.4f" % min(t(plan, 30) for _ in range(2)), flush=True)
for occ in ("0", "1", "2", "3", "4", "5"):
    for chunks in ("1", "2"):
        os.environ.update(AHMC_PIPE_UP="direct", AHMC_PIPE_DOWN="direct", AHMC_PIPE_CHUNKS=chunks, AHMC_PIPE_OCC=occ)
        print("direct/direct occ", occ, "chunks", chunks, "e2e call ms %.4f" % min(t(plan, 30) for _ in range(3)), flush=True)
        check("occ " + occ)
os.environ.pop("AHMC_PIPE_OCC")
for k in ("AHMC_PIPE_UP", "AHMC_PIPE_DOWN", "AHMC_PIPE_CHUNKS"): os.environ.pop(k)
print("library defaults: e2e call ms %.4f" % min(t(plan, 40) for _ in range(2)))
check("library defaults")
import time as _t
per = []
for _ in range(40):
    t0 = _t.perf_counter(); plan(); per.append((_t.perf_counter() - t0) * 1e3)
print("per-call wall ms: min %.4f med %.4f max %.4f" % (min(per), sorted(per)[20], max(per)), flush=True)
os.environ["AHMC_PIPE_TRACE"] = "1"
for _ in range(4): plan()
os.environ.pop("AHMC_PIPE_TRACE")
# how much of the bidirectional PCIe rate survives small copies?  n copies of `mb` MiB each way, two streams
for mb in (0.5, 1, 2, 4, 12):
    n = int(mb * 2 ** 20 // 8)
    reps = max(1, int(24 / mb))
    hs, ds = h_a[:n], d_a[:n]
    hd, dd = h_a2[:n], d_b[:n]
    def bid():
        with torch.cuda.stream(s1):
            for _ in range(reps): ds.copy_(hs, non_blocking=True)
        with torch.cuda.stream(s2):
            for _ in range(reps): hd.copy_(dd, non_blocking=True)
    def uni():
        with torch.cuda.stream(s1):
            for _ in range(reps): ds.copy_(hs, non_blocking=True)
    tb, tu = t(bid, 5), t(uni, 5)
    gb = reps * n * 8 / 1e9
    print("copies of %.1f MiB: H2D alone %.1f GB/s; H2D || D2H %.1f GB/s each way" % (mb, gb / (tu * 1e-3), gb / (tb * 1e-3)), flush=True)
# small batches: latency of one call
for n in (256, 1024):
    for k in ("AHMC_PIPE_UP", "AHMC_PIPE_DOWN", "AHMC_PIPE_CHUNKS", "AHMC_PIPE_TRACE"): os.environ.pop(k, None)
    zs = A.PhasePoint(thp.numpy()[:n], rp.numpy()[:n], A.DualValue(z0.lp.value[:n], gp.numpy()[:n]), A.DualValue(z0.lk.value[:n], None))
    zo = A.PhasePoint(outs[0].numpy()[:n], outs[1].numpy()[:n], A.DualValue(outs[3].numpy()[:n], outs[2].numpy()[:n]), A.DualValue(outs[4].numpy()[:n], None))
    pl = A.StepPlan(A.Leapfrog(0.1), h, zs, 32, out=zo)
    print("N", n, "defaults: call ms %.4f" % t(pl, 50))
    os.environ["AHMC_PIPE_UP"], os.environ["AHMC_PIPE_DOWN"] = "ce1", "ce"
    print("N", n, "ce1/ce:   call ms %.4f" % t(pl, 50))
