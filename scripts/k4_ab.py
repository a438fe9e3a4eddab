"""K4 tile-shape A/B: 16-chain tiles, two CTAs per SM (default) vs 32-chain tiles, one CTA per SM (AHMC_DENSE_TILE=32x1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ahmc_b200 as A
import bench

dev = torch.device("cuda:0")
for N in (4096, 16384):
    m, s, Minv, th, r = bench.synth(N, 128, 1)
    rng4 = np.random.Generator(np.random.PCG64(bench.SEED))
    Q, _ = np.linalg.qr(rng4.normal(size=(128, 128)))
    lam = np.exp(np.linspace(np.log(0.1), np.log(10.0), 128))
    hd = A.Hamiltonian(A.DiagEuclideanMetric(np.diag((Q * lam) @ Q.T).copy()), A.DenseGaussian(np.zeros(128), (Q / lam) @ Q.T))
    zd = A.phasepoint(hd, torch.as_tensor(th, device=dev), torch.as_tensor(r, device=dev))
    pd = A.StepPlan(A.Leapfrog(0.02), hd, zd, 32, flags=A.FLAG_ASYNC)
    ref = None
    for tile in ("", "32x1"):
        if tile:
            os.environ["AHMC_DENSE_TILE"] = tile
        else:
            os.environ.pop("AHMC_DENSE_TILE", None)
        for _ in range(3):
            z = pd()
        torch.cuda.synchronize()
        t0 = time.perf_counter()  # wall clock around back-to-back async launches (the context stream is not torch's)
        for _ in range(20):
            z = pd()
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 20 * 1e3
        th_out = z.theta.clone()
        if ref is None:
            ref = th_out
        print(f"N={N} tile={tile or '16x2'}: {ms:.4f} ms/trajectory, {2.0 * 128 * 128 * N * 32 / ms / 1e9:.2f} TFLOP/s, "
              f"max |dtheta| vs default {float((th_out - ref).abs().max()):.3e}", flush=True)
