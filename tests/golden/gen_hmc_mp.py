"""gen_hmc_mp.py -- known-answer fixtures for STATIC HMC transitions: tests/golden/hmc_mp50.json.

Independent restatement, in 50-digit mpmath arithmetic, of
  * `rand_momentum` (src/metric.jl:290-320: Unit r = z; Diag r = z ./ sqrt(M^-1); Dense r = U \\ z, U = chol(M^-1).U),
  * the static `transition` (src/trajectory.jl:271-300) with `EndPointTS` (`sample_phasepoint` :336-340,
    `mh_accept_ratio` :863-880, accept / revert :312-332, momentum flip :283) and
  * with `MultinomialTS` (:344-390: backward + forward trajectory, `randcat` by inverse CDF, src/utilities.jl:92-103,
    acceptance statistic = mean of min(1, exp(H0 - H_i)) over the whole trajectory),
from given standard normals, exponentials / uniforms and (for MultinomialTS) the forward-step count shared by all chains.
It shares no code with oracle/ahmc_oracle.c or the CUDA kernels.  Every accept / index comparison records its margin;
cases with a relative margin below 1e-7 are redrawn, so an fp64 implementation cannot legitimately decide otherwise.

Run:  python tests/golden/gen_hmc_mp.py      (rewrites hmc_mp50.json deterministically)
"""
import json
import os
import sys

import mpmath as mp
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_closed_form import dHdr_mp, logp_grad_mp, mpf_list  # noqa: E402

mp.mp.dps = 50
HERE = os.path.dirname(os.path.abspath(__file__))


def chol_upper(M, D):
    """U with U'U = M (LinearAlgebra.cholesky(Symmetric(M)).U)."""
    U = [[mp.mpf(0)] * D for _ in range(D)]
    for j in range(D):
        s = M[j][j] - sum(U[k][j] ** 2 for k in range(j))
        U[j][j] = mp.sqrt(s)
        for i in range(j + 1, D):
            U[j][i] = (M[j][i] - sum(U[k][j] * U[k][i] for k in range(j))) / U[j][j]
    return U


def rand_momentum(mkind, Minv, z, D):
    if mkind == "unit":
        return list(z)
    if mkind == "diag":
        return [z[d] / mp.sqrt(Minv[d]) for d in range(D)]
    U = chol_upper(Minv, D)  # back substitution U r = z
    r = [mp.mpf(0)] * D
    for i in reversed(range(D)):
        r[i] = (z[i] - sum(U[i][k] * r[k] for k in range(i + 1, D))) / U[i][i]
    return r


class Sys:
    def __init__(self, kind, D, p0, p1, c0, mkind, Minv, eps):
        self.kind, self.D, self.p0, self.p1, self.c0, self.mkind, self.Minv, self.eps = kind, D, p0, p1, c0, mkind, Minv, eps

    def point(self, th, r):
        lp, grad = logp_grad_mp(self.kind, self.D, self.p0, self.p1, self.c0, th)
        dr = dHdr_mp(self.mkind, self.Minv, r, self.D)
        return dict(th=th, r=r, lp=lp, g=[-x for x in grad], lk=-sum(a * b for a, b in zip(r, dr)) / 2)

    def step(self, z, fwd, i=1, n=1, alpha=None):  # leapfrog step i of an n-step `step` call (integrator.jl:233-247)
        e = self.eps if fwd else -self.eps
        D = self.D
        r0 = z["r"]
        if alpha is not None:  # TemperedLeapfrog (integrator.jl:198-209): i_temper = 2(i-1) + 1 for the first half
            sa = mp.sqrt(alpha)
            r0 = [x * sa for x in r0] if 2 * (i - 1) + 1 <= n else [x / sa for x in r0]
        r = [r0[d] - e / 2 * z["g"][d] for d in range(D)]
        dr = dHdr_mp(self.mkind, self.Minv, r, D)
        th = [z["th"][d] + e * dr[d] for d in range(D)]
        lp, grad = logp_grad_mp(self.kind, D, self.p0, self.p1, self.c0, th)
        g = [-x for x in grad]
        r = [r[d] - e / 2 * g[d] for d in range(D)]
        if alpha is not None:  # second half: i_temper = 2(i-1) + 2
            r = [x * sa for x in r] if 2 * (i - 1) + 2 <= n else [x / sa for x in r]
        dr = dHdr_mp(self.mkind, self.Minv, r, D)
        return dict(th=th, r=r, lp=lp, g=g, lk=-sum(a * b for a, b in zip(r, dr)) / 2)


def energy(z):
    return -(z["lp"] + z["lk"])


def make_case(rng, name, kind, mkind, D, N, eps, n_steps, ts, n_fwd=None, scale=1.0, temper=None):
    p0 = p1 = Minv = None
    c0 = 0.0
    if kind == "diag_gauss":
        p0, p1 = rng.normal(size=D), np.exp(rng.uniform(-0.5, 0.5, size=D))
    elif kind == "dense_gauss":
        A = rng.normal(size=(D, D))
        p0, p1 = rng.normal(size=D), A @ A.T / D + np.eye(D)
    if mkind == "diag":
        Minv = np.exp(rng.uniform(-0.5, 0.5, size=D))
    elif mkind == "dense":
        B = rng.normal(size=(D, D))
        Minv = B @ B.T / D + 0.5 * np.eye(D)
    pp0 = None if p0 is None else mpf_list(p0)
    pp1 = None if p1 is None else ([mpf_list(row) for row in p1] if kind == "dense_gauss" else mpf_list(p1))
    Mm = None if Minv is None else ([mpf_list(row) for row in Minv] if mkind == "dense" else mpf_list(Minv))
    S = Sys(kind, D, pp0, pp1, mp.mpf(c0), mkind, Mm, mp.mpf(float(eps)))
    al = None if temper is None else mp.mpf(float(temper))
    while True:
        theta = rng.normal(size=(N, D)) * scale
        normals = rng.normal(size=(N, D))
        var = rng.exponential(size=N) if ts == "endpoint" else rng.uniform(size=N)
        out = dict(theta=[], r=[], lp_gradient=[], lp_value=[], lk_value=[], is_accept=[], acceptance_rate=[],
                   hamiltonian_energy_error=[], index=[])
        margin = mp.mpf(1)
        for c in range(N):
            r0 = rand_momentum(mkind, Mm, mpf_list(normals[c]), D)
            z = S.point(mpf_list(theta[c]), r0)
            H0 = energy(z)
            if ts == "endpoint":
                z1 = z
                for i in range(1, n_steps + 1):
                    z1 = S.step(z1, True, i, n_steps, al)
                H1 = energy(z1)
                ex = mp.mpf(float(var[c]))
                margin = min(margin, abs(H1 - (H0 + ex)) / max(abs(H1), abs(H0 + ex), 1))
                acc = H1 < H0 + ex  # mh_accept_ratio (:863-867)
                alpha = min(mp.mpf(1), mp.e ** (H0 - H1))
                zn = z1 if acc else z
                idx = n_steps if acc else 0
            else:
                fwd, bwd = [], []
                zz = z
                for i in range(1, n_fwd + 1):  # each leg is its own `step` call: it tempers by its own number of steps
                    zz = S.step(zz, True, i, n_fwd, al)
                    fwd.append(zz)
                zz = z
                for i in range(1, n_steps - n_fwd + 1):
                    zz = S.step(zz, False, i, n_steps - n_fwd, al)
                    bwd.append(zz)
                zs = list(reversed(bwd)) + [z] + fwd  # :377
                lw = [-energy(q) for q in zs]
                m = max(lw)
                lse = m + mp.log(sum(mp.e ** (x - m) for x in lw))
                P = [mp.e ** (x - lse) for x in lw]
                u = mp.mpf(float(var[c]))
                cum, count = mp.mpf(0), 0
                for p in P:  # randcat (utilities.jl:92-103): count(C .< u) + 1
                    cum += p
                    margin = min(margin, abs(cum - u))
                    if cum < u:
                        count += 1
                i = min(max(count + 1, 1), len(zs))
                zn = zs[i - 1]
                idx = (i - 1) - (n_steps - n_fwd)  # signed offset of the drawn point from z
                acc = True
                alpha = sum(mp.e ** min(mp.mpf(0), -(energy(q) - H0)) for q in zs) / len(zs)  # :384-387
            Hn = energy(zn)
            out["theta"].append([float(x) for x in zn["th"]])
            out["r"].append([float(-x) for x in zn["r"]])  # momentum flip (:283)
            out["lp_gradient"].append([float(x) for x in zn["g"]])
            out["lp_value"].append(float(zn["lp"]))
            out["lk_value"].append(float(zn["lk"]))
            out["is_accept"].append(bool(acc))
            out["acceptance_rate"].append(float(alpha))
            out["hamiltonian_energy_error"].append(float(Hn - H0))
            out["index"].append(int(idx))
        if margin > mp.mpf("1e-7"):
            break
    tolist = lambda a: None if a is None else np.asarray(a).tolist()
    return dict(name=name, model=kind, metric=mkind, D=D, N=N, eps=float(eps), n_steps=n_steps, sampler=ts, n_fwd=n_fwd,
                temper_alpha=0.0 if temper is None else float(temper),
                p0=tolist(p0), p1=tolist(p1), c0=c0, Minv=tolist(Minv), theta0=theta.tolist(), normals=normals.tolist(),
                variates=var.tolist(), min_margin=float(margin), expect=out)


def main():
    rng = np.random.Generator(np.random.PCG64(20260925))
    cases = [
        make_case(rng, "ep_stdnormal_unit", "std_normal", "unit", 5, 8, 0.3, 8, "endpoint"),
        make_case(rng, "ep_diag_diag", "diag_gauss", "diag", 6, 8, 0.6, 6, "endpoint"),
        make_case(rng, "ep_dense_dense", "dense_gauss", "dense", 4, 8, 0.7, 5, "endpoint"),
        make_case(rng, "ep_funnel_diag_rejects", "funnel", "diag", 4, 10, 0.9, 6, "endpoint", scale=1.5),
        make_case(rng, "mn_diag_diag", "diag_gauss", "diag", 5, 8, 0.5, 7, "multinomial", n_fwd=3),
        make_case(rng, "mn_dense_unit_allfwd", "dense_gauss", "unit", 4, 6, 0.4, 5, "multinomial", n_fwd=5),
        make_case(rng, "mn_funnel_dense_allbwd", "funnel", "dense", 3, 6, 0.3, 6, "multinomial", n_fwd=0),
        # TemperedLeapfrog(eps, alpha) as the transition's integrator (appended: the cases above keep their random streams)
        make_case(rng, "ep_diag_diag_tempered_odd", "diag_gauss", "diag", 6, 8, 0.5, 7, "endpoint", temper=1.1),
        make_case(rng, "ep_funnel_unit_tempered", "funnel", "unit", 4, 8, 0.4, 6, "endpoint", temper=0.93),
        make_case(rng, "mn_diag_diag_tempered", "diag_gauss", "diag", 5, 8, 0.45, 7, "multinomial", n_fwd=3, temper=1.08),
    ]
    with open(os.path.join(HERE, "hmc_mp50.json"), "w") as f:
        json.dump(dict(generator="tests/golden/gen_hmc_mp.py", digits=50, cases=cases), f)
    for c in cases:
        e = c["expect"]
        print(f"{c['name']:28s} accept {sum(e['is_accept'])}/{c['N']} index {e['index']} min margin {c['min_margin']:.2e}")


if __name__ == "__main__":
    main()
