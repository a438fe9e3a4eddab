"""run_configs.py -- the five BASELINE.json configs end to end (sampling with adaptation where the config says so),
with throughput in leapfrog-steps*dims/s.  Single process = one GPU; under torchrun each rank takes its shard of
the chains and the adaptor record is all-gathered (NCCL).  Writes one JSON line per config.

  python scripts/run_configs.py [c1 c2 c3 c4 c5] [--scale 1.0] [--iters N]
  python -m torch.distributed.run --nproc-per-node 8 ... scripts/run_configs.py c4 c5
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ahmc_b200 as A  # noqa: E402
from ahmc_b200 import adaptation as ad  # noqa: E402

SEED = 20260923


def correlated_gaussian(D, seed):
    """Sigma = Q diag(lambda) Q', lambda log-spaced 1e-1..1e1, Q from QR of a seeded Gaussian."""
    rng = np.random.Generator(np.random.PCG64(seed))
    Q, _ = np.linalg.qr(rng.normal(size=(D, D)))
    lam = np.exp(np.linspace(np.log(0.1), np.log(10.0), D))
    Sigma = (Q * lam) @ Q.T
    P = (Q / lam) @ Q.T
    return Sigma, P


def run(name, world, rank, dev, scale, iters):
    import zlib

    rng = np.random.Generator(np.random.PCG64(SEED + zlib.crc32(name.encode()) % 1000 + rank))
    t_adapt = 0
    if name == "c1":   # StaticTrajectory(Leapfrog(0.1), 32) + Unit, D=10 std-normal, 64 chains
        D, N = 10, 64
        h = A.Hamiltonian(A.UnitEuclideanMetric(D), A.StdNormal(D))
        kern = A.HMCKernel(A.Trajectory(A.EndPointTS, A.Leapfrog(0.1), A.FixedNSteps(32)))
        adaptor, n_adapts, n_samples = None, 0, iters or 200
    elif name == "c2":  # HMCDA(0.8, lambda=1) + Diag, D=128 correlated Gaussian, 4096 chains
        D, N = 128, int(4096 * scale)
        Sigma, P = correlated_gaussian(D, SEED)
        h = A.Hamiltonian(A.DiagEuclideanMetric(np.diag(Sigma).copy()), A.DenseGaussian(np.zeros(D), P))
        kern = A.HMCKernel(A.Trajectory(A.EndPointTS, A.Leapfrog(0.1), A.FixedIntegrationTime(1.0)))
        adaptor = ad.NaiveHMCAdaptor(ad.UnitMassMatrix(), ad.NesterovDualAveraging(0.8, 0.1))
        n_samples = iters or 60
        n_adapts = n_samples // 2
    elif name == "c3":  # NUTS(Multinomial, GeneralisedNoUTurn) + Diag, D=128 Gaussian, 4096 chains
        D, N = 128, int(4096 * scale)
        s = np.exp(np.linspace(np.log(0.1), np.log(10.0), D))
        h = A.Hamiltonian(A.DiagEuclideanMetric(s * s), A.DiagGaussian(np.zeros(D), s))
        kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.1), A.GeneralisedNoUTurn()))
        adaptor, n_adapts, n_samples = None, 0, iters or 20
    elif name == "c4":  # NUTS + StanHMCAdaptor, funnel D=100, 32768 chains over the ranks
        D, N = 100, int(32768 * scale) // world
        h = A.Hamiltonian(A.DiagEuclideanMetric(np.ones(D)), A.Funnel(D))
        kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.1), A.GeneralisedNoUTurn()))
        adaptor = ad.StanHMCAdaptor(ad.WelfordVar(D), ad.NesterovDualAveraging(0.8, 0.1))
        n_samples = iters or 120
        n_adapts = int(n_samples * 0.8)
    elif name == "c4d":  # C4, pooled StanHMCAdaptor resident on the DEVICE (ahmc_adapt_exchange_f64 over an ahmc_comm): no host syncs
        D, N = 100, int(32768 * scale) // world
        h = A.Hamiltonian(A.DiagEuclideanMetric(np.ones(D)), A.Funnel(D))
        kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.1), A.GeneralisedNoUTurn()))
        adaptor = "device"
        n_samples = iters or 120
        n_adapts = int(n_samples * 0.8)
    elif name == "c4v":  # C4 with the reference's vectorised (per-chain) adaptors: warm-up + sampling in ONE launch
        D, N = 100, int(32768 * scale) // world
        h = A.Hamiltonian(A.DiagEuclideanMetric(np.ones(D)), A.Funnel(D))
        kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.1), A.GeneralisedNoUTurn()))
        adaptor = A.VectorisedStanAdaptor(delta=0.8)
        n_samples = iters or 120
        n_adapts = int(n_samples * 0.8)
    elif name == "c5":  # NUTS + DenseEuclideanMetric, D=256 Gaussian, 8192 chains over the ranks
        D, N = 256, int(8192 * scale) // world
        Sigma, P = correlated_gaussian(D, SEED + 5)
        h = A.Hamiltonian(A.DenseEuclideanMetric(Sigma), A.DenseGaussian(np.zeros(D), P))
        kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.5), A.GeneralisedNoUTurn()))
        adaptor, n_adapts, n_samples = None, 0, iters or 5
    else:
        raise SystemExit(f"unknown config {name}")
    theta0 = torch.as_tensor(rng.normal(size=(N, D)) * (0.1 if name.startswith("c4") else 1.0), device=dev)
    prng = A.PhiloxRNG(SEED + 17 * rank)
    # warm-up launch (module load, workspace allocation)
    ad.sample(prng, h, kern, theta0, 1)
    torch.cuda.synchronize()
    if adaptor == "device":
        comm = ad.Comm.from_torch_distributed(dev.index or 0) if world > 1 else None
        ad.sample_pooled_device(prng, h, kern, theta0, 3, 2, eps0=0.1, comm=comm)  # first NCCL call, buffers
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        t0 = time.perf_counter()
        res = ad.sample_pooled_device(prng, h, kern, theta0, n_samples, n_adapts, eps0=0.1, delta=0.8, comm=comm)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if comm is not None:
            comm.destroy()
    else:
        t0 = time.perf_counter()
        res = ad.sample(prng, h, kern, theta0, n_samples, adaptor, n_adapts, keep_draws=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    tt = torch.tensor([dt, float(res.leapfrog_steps)], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist

        tmax = tt.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        dt, steps = tmax[0].item(), tt[1].item()
    else:
        steps = res.leapfrog_steps
    th = res.theta
    out = dict(config=name, n_gpus=world, chains=N * world, D=D, transitions=n_samples, n_adapts=n_adapts,
               seconds=dt, leapfrog_steps=steps, rate_steps_dims_per_s=steps * D / dt,
               final_eps=float(res.eps) if np.ndim(res.eps) == 0 else float(res.eps.double().median().item()),
               mean_accept=float(np.mean([s["acceptance_rate"] for s in res.stats[n_adapts:]])),
               divergent=int(sum(s["numerical_error"] for s in res.stats[n_adapts:])),
               theta_mean_abs_max=float(th.mean(dim=0).abs().max().item()), theta_std_first=float(th[:, 0].std().item()),
               timing="wall clock around the whole sample() loop (python host loop + kernels + adaptor exchange)",
               breakdown_s={k: round(v, 4) for k, v in res.timing.items()})
    if rank == 0:
        print(json.dumps(out), flush=True)


def main():
    argv = sys.argv[1:]
    args = [a for i, a in enumerate(argv) if not a.startswith("--") and (i == 0 or argv[i - 1] not in ("--scale", "--iters"))]
    scale = float(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 1.0
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 0
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    for name in args or ["c1", "c2", "c3", "c4", "c5"]:
        run(name, world, rank, dev, scale, iters)
    if world > 1:
        import torch.distributed as dist

        dist.destroy_process_group()


if __name__ == "__main__":
    main()
