"""adaptation.py -- host-side mirror of src/adaptation/*.jl for MANY chains with POOLED statistics, and the
`sample` loop (src/sampler.jl:159-248) on top of the fused transition kernels.

The reference adapts per chain and never pools (`src/adaptation/Adaptation.jl:52`: "TODO: implement consensus
adaptor"); many-chain HMCDA / NUTS with one shared step size and metric needs the pooled form (SURVEY.md 8a Q6,
8e).  Each GPU reduces its chains to a (2+2D)-double record with ahmc_adapt_summary_f64 (K5); the records of all
ranks are exchanged with ONE all-gather (torch.distributed / NCCL) and merged in rank order, so every rank computes
bit-identical eps and M^-1.  With one chain on one rank every formula below reduces to the reference's scalar path
(checked against the oracle in tests/test_adaptation.py).  Scalar arithmetic only: this is control logic, not the
hot path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import core as A


# ------------------------------------------------------------------------------------------------ records
def merge_records(records: List[np.ndarray]) -> np.ndarray:
    """Chan merge of per-rank records [n, sum_alpha, mean[D], M2[D]] in list (= rank) order."""
    out = np.array(records[0], dtype=np.float64, copy=True)
    D = (out.size - 2) // 2
    for rec in records[1:]:
        n_a, n_b = out[0], rec[0]
        n = n_a + n_b
        delta = rec[2:2 + D] - out[2:2 + D]
        out[2 + D:] = out[2 + D:] + rec[2 + D:] + delta * delta * (n_a * n_b / n)
        out[2:2 + D] = out[2:2 + D] + delta * (n_b / n)
        out[1] = out[1] + rec[1]
        out[0] = n
    return out


def allgather_records(rec) -> List[np.ndarray]:
    """One all-gather of the tiny adaptor record across ranks (the path's only exchange)."""
    try:
        import torch
        import torch.distributed as dist
    except Exception:  # pragma: no cover
        dist = None
    if dist is None or not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        r = rec.detach().cpu().numpy() if hasattr(rec, "detach") else np.asarray(rec)
        return [np.array(r, dtype=np.float64)]
    t = rec if hasattr(rec, "detach") else torch.as_tensor(rec)
    if dist.get_backend() == "nccl" and not t.is_cuda:
        t = t.cuda()
    if dist.get_backend() == "gloo" and t.is_cuda:
        t = t.cpu()
    outs = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, t.contiguous())
    return [o.cpu().numpy().astype(np.float64) for o in outs]


# ------------------------------------------------------------------------------------------------ step size
class NesterovDualAveraging:
    """src/adaptation/stepsize.jl:111-229 with a scalar (shared) step size; `adapt` takes the pooled mean of
    min(1, alpha) over all chains (one chain: exactly stepsize.jl:178-210)."""

    def __init__(self, delta: float, eps: float, gamma: float = 0.05, t_0: float = 10.0, kappa: float = 0.75):
        self.gamma, self.t_0, self.kappa, self.delta = gamma, t_0, kappa, delta
        self.eps = float(eps)
        self.reset()

    def reset(self):  # stepsize.jl:38-44
        self.m, self.mu, self.x_bar, self.H_bar = 0, math.log(10 * self.eps), 0.0, 0.0

    def adapt(self, mean_min1_alpha: float):
        m = self.m + 1
        eta_H = 1.0 / (m + self.t_0)
        H_bar = (1.0 - eta_H) * self.H_bar + eta_H * (self.delta - mean_min1_alpha)
        x = self.mu - H_bar * (math.sqrt(m) / self.gamma)
        eta_x = m ** (-self.kappa)
        x_bar = (1.0 - eta_x) * self.x_bar + eta_x * x
        try:
            eps = math.exp(x)
        except OverflowError:
            eps = math.inf
        if not math.isfinite(eps):  # stepsize.jl:199-203: keep the previous state
            return
        self.m, self.eps, self.x_bar, self.H_bar = m, eps, x_bar, H_bar

    def finalize(self):  # stepsize.jl:54-57
        self.eps = math.exp(self.x_bar)


class FixedStepSize:
    def __init__(self, eps):
        self.eps = eps

    def adapt(self, a): ...
    def reset(self): ...
    def finalize(self): ...


# ------------------------------------------------------------------------------------------------ mass matrix
class WelfordVar:
    """Pooled WelfordVar (src/adaptation/massmatrix.jl:86-157): every chain of every rank contributes one sample
    per iteration; batches arrive as (n, mean, M2) records and are Chan-merged."""

    def __init__(self, D: int, n_min: int = 10):
        self.D, self.n_min = D, n_min
        self.var = np.ones(D)
        self.reset()

    def reset(self):
        self.n, self.mu, self.M = 0.0, np.zeros(self.D), np.zeros(self.D)

    def push_record(self, rec: np.ndarray):
        D = self.D
        n_b, mean_b, M2_b = rec[0], rec[2:2 + D], rec[2 + D:2 + 2 * D]
        n = self.n + n_b
        delta = mean_b - self.mu
        self.M = self.M + M2_b + delta * delta * (self.n * n_b / n)
        self.mu = self.mu + delta * (n_b / n)
        self.n = n

    def get_estimation(self):  # massmatrix.jl:152-157
        n = self.n
        return n / ((n + 5) * (n - 1)) * self.M + 1e-3 * (5 / (n + 5))

    def update(self):  # massmatrix.jl:60-62
        if self.n >= self.n_min:
            self.var = self.get_estimation()


class UnitMassMatrix:
    var = None

    def reset(self): ...
    def push_record(self, rec): ...
    def update(self): ...


# ------------------------------------------------------------------------------------------------ Stan windows
def stan_windows(n_adapts: int, init_buffer: int = 75, term_buffer: int = 50, window_size: int = 25):
    """src/adaptation/stan_adaptor.jl:13-50 -> (window_start, window_end, window_splits)."""
    window_start, window_end = init_buffer + 1, n_adapts - term_buffer
    splits, next_window = [], init_buffer + window_size
    while next_window <= window_end:
        if next_window + 2 * window_size > window_end:
            next_window = window_end
        splits.append(next_window)
        window_size *= 2
        next_window += window_size
    if splits and splits[-1] == n_adapts:
        splits.pop()
    return window_start, window_end, splits


class StanHMCAdaptor:
    """src/adaptation/stan_adaptor.jl:61-159 (3-phase windowed adaptation) on pooled records."""

    def __init__(self, pc, ssa, init_buffer: int = 75, term_buffer: int = 50, window_size: int = 25):
        self.pc, self.ssa = pc, ssa
        self.init_buffer, self.term_buffer, self.window_size = init_buffer, term_buffer, window_size
        self.i, self.window_start, self.window_end, self.window_splits = 0, 0, 0, []

    def initialize(self, n_adapts: int):
        self.window_start, self.window_end, self.window_splits = stan_windows(
            n_adapts, self.init_buffer, self.term_buffer, self.window_size)

    def adapt(self, rec: np.ndarray):
        """rec = merged record of this iteration (stan_adaptor.jl:137-159)."""
        self.i += 1
        self.ssa.adapt(rec[1] / rec[0])
        if self.window_start <= self.i <= self.window_end:
            self.pc.push_record(rec)
            if self.i in self.window_splits:
                self.pc.update()
        if self.i in self.window_splits:
            self.ssa.reset()
            self.pc.reset()

    def finalize(self):
        self.ssa.finalize()

    @property
    def eps(self):
        return self.ssa.eps

    @property
    def Minv(self):
        return self.pc.var


class NaiveHMCAdaptor(StanHMCAdaptor):
    """src/adaptation/Adaptation.jl:41-64: adapt both every iteration, no windows."""

    def initialize(self, n_adapts: int): ...

    def adapt(self, rec: np.ndarray):
        self.i += 1
        self.ssa.adapt(rec[1] / rec[0])
        self.pc.push_record(rec)
        self.pc.update()


# ------------------------------------------------------------------------------------------------ sample
@dataclass
class SampleResult:
    theta: object          # final positions (N, D)
    draws: list            # kept draws (list of (N, D) tensors) if keep_draws
    stats: list            # per-iteration dicts of pooled scalars
    eps: float
    Minv: Optional[np.ndarray]
    leapfrog_steps: int = 0


def sample(rng, h: A.Hamiltonian, kappa: A.HMCKernel, theta, n_samples: int, adaptor=None, n_adapts: int = 0,
           keep_draws: bool = False, drop_warmup: bool = False) -> SampleResult:
    """`sample(rng, h, kappa, theta, n_samples, adaptor, n_adapts)` (src/sampler.jl:159-248) for N chains on this
    rank, pooled adaptation across chains and ranks.  Per iteration: one fused transition kernel (K2 / K3), and
    during warm-up one K5 launch + one all-gather of the (2+2D)-double record."""
    import torch

    z = A.phasepoint(h, theta, torch.zeros_like(theta))  # sample_init (sampler.jl:36-46); r is refreshed anyway
    draws, stats, total_steps = [], [], 0
    if adaptor is not None and n_adapts > 0:
        adaptor.initialize(n_adapts)
    for i in range(1, n_samples + 1):
        tr = A.transition(rng, h, kappa, z)
        z = tr.z
        if adaptor is not None and i <= n_adapts:  # Adaptation.adapt! glue (sampler.jl:72-90)
            rec = merge_records(allgather_records(A.adapt_summary(z.theta, tr.stat["acceptance_rate"])))
            adaptor.adapt(rec)
            if i == n_adapts:
                adaptor.finalize()
            if adaptor.Minv is not None:
                h = A.Hamiltonian(A.renew(h.metric, np.array(adaptor.Minv)), h.target)  # update(h, adaptor)
            tau = kappa.tau
            kappa = A.HMCKernel(A.Trajectory(tau.sampler, A.update_nom_step_size(tau.integrator, adaptor.eps),
                                             tau.termination_criterion), kappa.refreshment)
        ns = tr.stat["n_steps"]
        total_steps += int(ns.sum().item()) if hasattr(ns, "sum") else int(ns) * theta.shape[0]
        stats.append(dict(acceptance_rate=float(tr.stat["acceptance_rate"].mean().item()),
                          step_size=A.step_size(kappa.tau.integrator),
                          numerical_error=int(tr.stat["numerical_error"].sum().item()), is_adapt=i <= n_adapts))
        if keep_draws and (not drop_warmup or i > n_adapts):
            draws.append(z.theta.clone())
    return SampleResult(z.theta, draws, stats, A.step_size(kappa.tau.integrator),
                        None if adaptor is None else adaptor.Minv, total_steps)
