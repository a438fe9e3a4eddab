"""gen_adapt_mp.py -- known-answer fixtures for the adaptors: tests/golden/adapt_mp50.json.

50-digit mpmath restatement, written from the Julia source, of
  * `NesterovDualAveraging`: `DAState` / `adapt_stepsize!` / `reset!` / `finalize!` (src/adaptation/stepsize.jl:25-62,
    178-210) on a sequence of acceptance rates (some > 1, exercising min(1, alpha)), with a reset in the middle;
  * `WelfordVar` `push!` / `get_estimation` (src/adaptation/massmatrix.jl:141-157) and `WelfordCov` (:324-340) on a
    sequence of positions;
  * `NutpieVar` `get_estimation` (:244-248) on positions + gradients.
Independent of oracle/ahmc_oracle.c and of advancedhmc.jl_b200/adaptation.py.

Run:  python tests/golden/gen_adapt_mp.py
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 50
HERE = os.path.dirname(os.path.abspath(__file__))


def da_run(eps0, alphas, reset_at, delta=0.8, gamma=0.05, t0=10.0, kappa=0.75):
    delta, gamma, t0, kappa = (mp.mpf(x) for x in (delta, gamma, t0, kappa))
    eps = mp.mpf(eps0)
    mu, m, x_bar, H_bar = mp.log(10 * eps), 0, mp.mpf(0), mp.mpf(0)  # DAState(eps) (:27-30)
    trace = []
    for i, a in enumerate(alphas, 1):
        m += 1
        eta_H = 1 / (m + t0)
        H_bar = (1 - eta_H) * H_bar + eta_H * (delta - min(mp.mpf(1), mp.mpf(float(a))))
        x = mu - H_bar * (mp.sqrt(m) / gamma)
        eta_x = mp.mpf(m) ** (-kappa)
        x_bar = (1 - eta_x) * x_bar + eta_x * x
        eps = mp.e ** x
        if i == reset_at:  # reset! (:38-44)
            m, mu, x_bar, H_bar = 0, mp.log(10 * eps), mp.mpf(0), mp.mpf(0)
        trace.append(float(eps))
    return trace, float(mp.e ** x_bar)  # finalize! (:54-57)


def welford_var(xs):
    D = xs.shape[1]
    n, mu, M = 0, [mp.mpf(0)] * D, [mp.mpf(0)] * D
    for s in xs:
        n += 1
        s = [mp.mpf(float(v)) for v in s]
        d = [s[k] - mu[k] for k in range(D)]
        mu = [mu[k] + d[k] / n for k in range(D)]
        M = [M[k] + d[k] * d[k] * (mp.mpf(n - 1) / n) for k in range(D)]
    est = [mp.mpf(n) / ((n + 5) * (n - 1)) * M[k] + mp.mpf("1e-3") * (mp.mpf(5) / (n + 5)) for k in range(D)]
    return n, mu, M, est


def welford_cov(xs):
    D = xs.shape[1]
    n, mu, M = 0, [mp.mpf(0)] * D, [[mp.mpf(0)] * D for _ in range(D)]
    for s in xs:
        n += 1
        s = [mp.mpf(float(v)) for v in s]
        d = [s[k] - mu[k] for k in range(D)]
        mu = [mu[k] + d[k] / n for k in range(D)]
        M = [[M[i][j] + (s[i] - mu[i]) * d[j] for j in range(D)] for i in range(D)]
    est = [[mp.mpf(n) / ((n + 5) * (n - 1)) * M[i][j] + (mp.mpf("1e-3") * (mp.mpf(5) / (n + 5)) if i == j else 0)
            for j in range(D)] for i in range(D)]
    return n, mu, M, est


def main():
    rng = np.random.Generator(np.random.PCG64(20260926))
    alphas = rng.uniform(0.2, 1.3, size=40)
    trace, final = da_run(0.1, alphas, reset_at=25)
    D = 4
    L = rng.normal(size=(D, D))
    xs = rng.normal(size=(30, D)) @ L + rng.normal(size=D)
    gs = -(xs - xs.mean(axis=0)) @ np.linalg.inv(L.T @ L)
    n, mu, M, est = welford_var(xs)
    _, _, _, est_g = welford_var(gs)
    _, mu_c, M_c, est_c = welford_cov(xs)
    f = lambda v: [float(a) for a in v]
    out = dict(generator="tests/golden/gen_adapt_mp.py", digits=50,
               dual_averaging=dict(eps0=0.1, delta=0.8, gamma=0.05, t0=10.0, kappa=0.75, alphas=alphas.tolist(), reset_at=25,
                                   eps_trace=trace, eps_final=final),
               welford=dict(xs=xs.tolist(), gs=gs.tolist(), n=n, mu=f(mu), M=f(M), var_estimate=f(est),
                            cov_M=[f(r) for r in M_c], cov_estimate=[f(r) for r in est_c],
                            nutpie_estimate=[float(mp.sqrt(a / b)) for a, b in zip(est, est_g)]))
    with open(os.path.join(HERE, "adapt_mp50.json"), "w") as fh:
        json.dump(out, fh)
    print("eps trace tail", trace[-3:], "final", final)


if __name__ == "__main__":
    main()
