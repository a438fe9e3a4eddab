"""oracle_np.py -- numpy twin of the reference's MATRIX-MODE path, op-for-op with the same temporaries.

*** TEST / BENCH INFRASTRUCTURE ONLY (lives under oracle/): never imported by the product package. ***

Every array expression of the reference's `step` (src/integrator.jl:216-265), `dH/dr` / `neg_energy`
(src/hamiltonian.jl:50-68,155-177), `PhasePoint` (-Inf mapping, :95-104) and the vectorised static
`transition` (src/trajectory.jl:271-340,863-880) is restated as ONE numpy expression that materialises
the same temporaries a Julia broadcast would (Julia does not fuse `r - eps/2 .* g`: the `.*` makes a
temporary, the undotted `-` another).  It serves two purposes:
  1. an independent second restatement that cross-checks oracle/ahmc_oracle.c (tests/test_oracle.py);
  2. the stand-in for "the reference's own single-threaded CPU vectorised path" in bench.py's
     cpu_baseline (the Julia reference cannot run here: no julia binary).
PARITY PIN STATUS: unpinned against the reference itself (see oracle/ahmc_oracle.h).

Arrays are (D, N) Fortran-ordered float64 (Julia's column-major D x N).
"""
from __future__ import annotations

import numpy as np

STD_NORMAL, DIAG_GAUSS, DENSE_GAUSS, FUNNEL = 0, 1, 2, 3
UNIT, DIAG, DENSE = 0, 1, 2


class Model:
    """The user closures l_pi / dl_pi/dtheta handed to `Hamiltonian` (src/hamiltonian.jl:1-6)."""

    def __init__(self, kind, D, p0=None, p1=None, c0=0.0):
        self.kind, self.D, self.c0 = kind, D, float(c0)
        self.p0 = None if p0 is None else np.asarray(p0, dtype=np.float64)
        self.p1 = None if p1 is None else np.asarray(p1, dtype=np.float64)

    def logp_grad(self, th):
        """returns (lp[N], grad[D,N]) -- what `h.dlp/dth(theta)` returns (hamiltonian.jl:46)."""
        if self.kind == STD_NORMAL:
            v = -(np.abs(th) ** 2) / 2
            return v.sum(axis=0) + self.c0, -th
        if self.kind == DIAG_GAUSS:  # test/common.jl:40-56 (true gradient)
            m, s = self.p0[:, None], self.p1[:, None]
            g = m - th
            v = -((np.abs(g) ** 2) / s**2) / 2
            return v.sum(axis=0) + self.c0, g / s**2
        if self.kind == DENSE_GAUSS:
            mu, P = self.p0[:, None], self.p1
            diff = th - mu
            Pd = P @ diff
            return -(diff * Pd).sum(axis=0) / 2 + self.c0, -Pd
        if self.kind == FUNNEL:
            v = th[0]
            ev = np.exp(-v)
            x = th[1:]
            S = (x * x * ev).sum(axis=0)
            lp = -(v * v) / 18 - (S + (self.D - 1) * v) / 2
            grad = np.empty_like(th)
            grad[0] = -v / 9 + (S - (self.D - 1)) / 2
            grad[1:] = -x * ev
            return lp + self.c0, grad
        raise ValueError(self.kind)


class Metric:
    def __init__(self, kind, Minv=None):
        self.kind = kind
        self.Minv = None if Minv is None else np.asarray(Minv, dtype=np.float64)
        if kind == DENSE:
            self.cholU = np.linalg.cholesky(self.Minv).T  # cholesky(Symmetric(Minv)).U  (metric.jl:108)

    def _Mb(self):  # D-vector broadcasts over columns; D x N used as is (metric.jl:64)
        return self.Minv[:, None] if self.Minv.ndim == 1 else self.Minv

    def dHdr(self, r):  # hamiltonian.jl:50-68
        if self.kind == UNIT:
            return r.copy()
        if self.kind == DIAG:
            return self._Mb() * r
        return self.Minv @ r

    def neg_energy(self, r):  # hamiltonian.jl:155-184
        if self.kind == UNIT:
            return -(np.abs(r) ** 2).sum(axis=0) / 2
        if self.kind == DIAG:
            return -((np.abs(r) ** 2) * self._Mb()).sum(axis=0) / 2
        return -(r * (self.Minv @ r)).sum(axis=0) / 2

    def rand_momentum(self, z):  # metric.jl:290-320
        if self.kind == UNIT:
            return z.copy()
        if self.kind == DIAG:
            return z / np.sqrt(self._Mb())
        import scipy.linalg

        return scipy.linalg.solve_triangular(self.cholU, z, lower=False)


def _map_nonfinite(v):  # hamiltonian.jl:95-104
    return np.where(np.isfinite(v), v, -np.inf)


class PhasePoint:
    __slots__ = ("theta", "r", "lp_value", "lp_gradient", "lk_value", "lk_gradient")

    def __init__(self, theta, r, lp_value, lp_gradient, lk_value, lk_gradient):
        self.theta, self.r = theta, r
        self.lp_value, self.lp_gradient = _map_nonfinite(lp_value), lp_gradient
        self.lk_value, self.lk_gradient = _map_nonfinite(lk_value), lk_gradient

    def isfinite(self):  # hamiltonian.jl:141-142  -- ALL chains at once (quirk Q1)
        return bool(
            np.all(np.isfinite(self.lp_value))
            and np.all(np.isfinite(self.lp_gradient))
            and np.all(np.isfinite(self.lk_value))
            and np.all(np.isfinite(self.lk_gradient))
        )

    def energy(self):  # hamiltonian.jl:149,194
        return -(self.lp_value + self.lk_value)


def phasepoint(model, metric, theta, r, lp=None):
    """hamiltonian.jl:115-119; lp=(value, MINUS gradient) if cached."""
    if lp is None:
        v, g = model.logp_grad(theta)
        lp = (v, -g)  # dH/dtheta: DualValue(res[1], -res[2])  (hamiltonian.jl:45-48)
    return PhasePoint(theta, r, lp[0], lp[1], metric.neg_energy(r), metric.dHdr(r))


def step(model, metric, eps, z, n_steps=1, temper_alpha=None, full_trajectory=False):
    """`step(lf, h, z, n_steps)` (integrator.jl:216-265), matrix mode, all-chains break."""
    fwd = n_steps > 0
    n_steps = abs(n_steps)
    eps = np.asarray(eps, dtype=np.float64)
    eps = eps if fwd else -eps
    if eps.ndim == 1:
        eps = eps[None, :]  # eps' : row vector broadcasting over columns (integrator.jl:227)
    res = []
    theta, r = z.theta, z.r
    value, gradient = z.lp_value, z.lp_gradient
    for i in range(1, n_steps + 1):
        if temper_alpha is not None:  # integrator.jl:198-209
            it = 2 * (i - 1) + 1
            r = r * np.sqrt(temper_alpha) if it <= n_steps else r / np.sqrt(temper_alpha)
        t1 = eps / 2 * gradient          # temporary:  eps/2 .* gradient
        r = r - t1                       # temporary:  r - (...)
        dr = metric.dHdr(r)              # temporary
        t2 = eps * dr                    # temporary
        theta = theta + t2               # temporary
        v, g = model.logp_grad(theta)    # user closure (its own temporaries)
        value, gradient = v, -g          # negate: temporary
        t3 = eps / 2 * gradient
        r = r - t3
        if temper_alpha is not None:
            it = 2 * (i - 1) + 2
            r = r * np.sqrt(temper_alpha) if it <= n_steps else r / np.sqrt(temper_alpha)
        z = phasepoint(model, metric, theta, r, lp=(value, gradient))  # neg_energy + 2nd dH/dr
        if full_trajectory:
            res.append(z)
        if not z.isfinite():
            break
    return res if full_trajectory else z


def hmc_transition(model, metric, eps, n_steps, z, normal_tape, exp_tape):
    """sampler.jl:48-58 (full refresh) + trajectory.jl:271-300 (EndPointTS), matrix mode."""
    r = metric.rand_momentum(normal_tape) if normal_tape is not None else z.r
    z0 = phasepoint(model, metric, z.theta, r)
    H0 = z0.energy()
    z1 = step(model, metric, eps, z0, n_steps)
    H1 = z1.energy()
    accept = H1 < H0 + exp_tape                       # trajectory.jl:869-877
    with np.errstate(over="ignore", invalid="ignore"):
        alpha = np.minimum(1.0, np.exp(H0 - H1))
    rej = ~accept
    th, rr, lg = z1.theta.copy(), z1.r.copy(), z1.lp_gradient.copy()
    lv, kv, kg = z1.lp_value.copy(), z1.lk_value.copy(), z1.lk_gradient.copy()
    th[:, rej], rr[:, rej], lg[:, rej] = z0.theta[:, rej], z0.r[:, rej], z0.lp_gradient[:, rej]
    lv[rej], kv[rej], kg[:, rej] = z0.lp_value[rej], z0.lk_value[rej], z0.lk_gradient[:, rej]
    znew = PhasePoint(th, -rr, lv, lg, kv, kg)        # flip (trajectory.jl:283)
    H = znew.energy()
    stat = dict(
        n_steps=n_steps, is_accept=accept, acceptance_rate=alpha, log_density=znew.lp_value,
        hamiltonian_energy=H, hamiltonian_energy_error=H - H0, numerical_error=not np.all(np.isfinite(H1)),
    )
    return znew, stat
