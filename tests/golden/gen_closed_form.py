"""gen_closed_form.py -- generate the committed known-answer fixtures tests/golden/*.json.

The reference (pure Julia) cannot run in this image and its tests hold no golden vectors for the
leapfrog path, so the known answers are produced independently of BOTH the
C oracle and the CUDA kernels: the leapfrog map of src/integrator.jl:233-247 is iterated in 50-digit
mpmath arithmetic (for Gaussian targets this IS the closed-form linear map A^n of SURVEY 8c, with
no rounding at the 1e-10 level), then rounded once to float64.

Run:  python tests/golden/gen_closed_form.py      (rewrites leapfrog_mp50.json deterministically)
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 50
HERE = os.path.dirname(os.path.abspath(__file__))


def mpf_list(a):
    return [mp.mpf(float(x)) for x in a]


def logp_grad_mp(kind, D, p0, p1, c0, th):
    """returns (lp, grad log pi) in mp arithmetic."""
    if kind == "std_normal":
        return c0 - sum(t * t for t in th) / 2, [-t for t in th]
    if kind == "diag_gauss":
        g = [p0[d] - th[d] for d in range(D)]
        return c0 - sum(g[d] * g[d] / (p1[d] * p1[d]) for d in range(D)) / 2, [g[d] / (p1[d] * p1[d]) for d in range(D)]
    if kind == "dense_gauss":
        diff = [th[d] - p0[d] for d in range(D)]
        Pd = [sum(p1[d][k] * diff[k] for k in range(D)) for d in range(D)]
        return c0 - sum(diff[d] * Pd[d] for d in range(D)) / 2, [-x for x in Pd]
    if kind == "funnel":
        v = th[0]
        ev = mp.e ** (-v)
        S = sum(th[d] * th[d] * ev for d in range(1, D))
        lp = c0 - v * v / 18 - (S + (D - 1) * v) / 2
        return lp, [-v / 9 + (S - (D - 1)) / 2] + [-th[d] * ev for d in range(1, D)]
    raise ValueError(kind)


def dHdr_mp(mkind, Minv, r, D):
    if mkind == "unit":
        return list(r)
    if mkind == "diag":
        return [Minv[d] * r[d] for d in range(D)]
    return [sum(Minv[d][k] * r[k] for k in range(D)) for d in range(D)]


def leapfrog_mp(kind, D, p0, p1, c0, mkind, Minv, eps, n_steps, th, r, temper_alpha=None):
    fwd = n_steps > 0
    n = abs(n_steps)
    eps = eps if fwd else -eps
    lp, grad = logp_grad_mp(kind, D, p0, p1, c0, th)
    g = [-x for x in grad]
    for i in range(1, n + 1):
        if temper_alpha is not None:
            sa = mp.sqrt(temper_alpha)
            r = [x * sa if 2 * (i - 1) + 1 <= n else x / sa for x in r]
        r = [r[d] - eps / 2 * g[d] for d in range(D)]
        dr = dHdr_mp(mkind, Minv, r, D)
        th = [th[d] + eps * dr[d] for d in range(D)]
        lp, grad = logp_grad_mp(kind, D, p0, p1, c0, th)
        g = [-x for x in grad]
        r = [r[d] - eps / 2 * g[d] for d in range(D)]
        if temper_alpha is not None:
            sa = mp.sqrt(temper_alpha)
            r = [x * sa if 2 * (i - 1) + 2 <= n else x / sa for x in r]
    dr = dHdr_mp(mkind, Minv, r, D)
    lk = -sum(r[d] * dr[d] for d in range(D)) / 2
    return th, r, g, lp, lk


def make_case(rng, name, kind, mkind, D, N, eps, n_steps, per_chain_eps=False, per_chain_minv=False,
              temper_alpha=None):
    p0 = p1 = None
    c0 = 0.0
    if kind == "diag_gauss":
        p0 = rng.normal(size=D)
        p1 = np.exp(rng.uniform(-1, 1, size=D))
        c0 = float(-0.5 * np.sum(np.log(2 * np.pi) + 2 * np.log(p1)))
    elif kind == "dense_gauss":
        p0 = rng.normal(size=D)
        A = rng.normal(size=(D, D))
        p1 = A @ A.T / D + np.eye(D)
    elif kind == "funnel":
        c0 = 0.0
    Minv = None
    if mkind == "diag":
        Minv = np.exp(rng.uniform(-1, 1, size=(D, N) if per_chain_minv else D))
    elif mkind == "dense":
        B = rng.normal(size=(D, D))
        Minv = B @ B.T / D + 0.5 * np.eye(D)
    theta = rng.normal(size=(D, N)) * (0.5 if kind == "funnel" else 1.0)
    r = rng.normal(size=(D, N))
    eps_arr = eps * np.exp(rng.uniform(-0.3, 0.3, size=N)) if per_chain_eps else None
    out = dict(theta=[], r=[], lp_gradient=[], lp_value=[], lk_value=[])
    for c in range(N):
        if mkind == "diag":
            Mc = mpf_list(Minv[:, c] if per_chain_minv else Minv)
        elif mkind == "dense":
            Mc = [mpf_list(row) for row in Minv]
        else:
            Mc = None
        pp0 = None if p0 is None else mpf_list(p0)
        pp1 = None if p1 is None else ([mpf_list(row) for row in p1] if kind == "dense_gauss" else mpf_list(p1))
        e = mp.mpf(float(eps_arr[c])) if per_chain_eps else mp.mpf(float(eps))
        th, rr, g, lp, lk = leapfrog_mp(kind, D, pp0, pp1, mp.mpf(c0), mkind, Mc, e, n_steps,
                                        mpf_list(theta[:, c]), mpf_list(r[:, c]),
                                        None if temper_alpha is None else mp.mpf(temper_alpha))
        out["theta"].append([float(x) for x in th])
        out["r"].append([float(x) for x in rr])
        out["lp_gradient"].append([float(x) for x in g])
        out["lp_value"].append(float(lp))
        out["lk_value"].append(float(lk))
    tolist = lambda a: None if a is None else np.asarray(a).tolist()
    return dict(name=name, model=kind, metric=mkind, D=D, N=N, eps=float(eps), eps_chain=tolist(eps_arr),
                n_steps=n_steps, temper_alpha=temper_alpha, p0=tolist(p0), p1=tolist(p1), c0=c0,
                Minv=tolist(Minv), theta0=theta.T.tolist(), r0=r.T.tolist(), expect=out)


def spot_values():
    """The three spot vectors recorded in SURVEY.md section 8c (1-D, eps = double nearest 0.1)."""
    res = []
    for (th0, r0, n, Minv, m, s) in [(1.0, 0.5, 32, 1.0, 0.0, 1.0), (1.0, 0.5, 32, 0.25, 0.5, 2.0),
                                     (-0.75, 1.25, 10, 1.0, 0.0, 1.0)]:
        th, r, g, lp, lk = leapfrog_mp("diag_gauss", 1, [mp.mpf(m)], [mp.mpf(s)], mp.mpf(0), "diag", [mp.mpf(Minv)],
                                       mp.mpf(0.1), n, [mp.mpf(th0)], [mp.mpf(r0)])
        res.append(dict(theta0=th0, r0=r0, n=n, Minv=Minv, m=m, s=s, theta=mp.nstr(th[0], 20), r=mp.nstr(r[0], 20)))
    return res


def main():
    rng = np.random.Generator(np.random.PCG64(20260923))
    cases = [
        make_case(rng, "c1_stdnormal_unit", "std_normal", "unit", 10, 4, 0.1, 32),
        make_case(rng, "stdnormal_unit_bwd", "std_normal", "unit", 5, 3, 0.1, -7),
        make_case(rng, "diag_diag", "diag_gauss", "diag", 7, 3, 0.1, 32),
        make_case(rng, "diag_diag_perchain_eps", "diag_gauss", "diag", 5, 4, 0.08, 10, per_chain_eps=True),
        make_case(rng, "diag_diag_perchain_minv", "diag_gauss", "diag", 6, 3, 0.05, 20, per_chain_minv=True),
        make_case(rng, "diag_unit_1step", "diag_gauss", "unit", 33, 2, 0.2, 1),
        make_case(rng, "diag_diag_tempered", "diag_gauss", "diag", 4, 2, 0.1, 6, temper_alpha=1.05),
        make_case(rng, "dense_diag", "dense_gauss", "diag", 6, 3, 0.1, 16),
        make_case(rng, "dense_dense", "dense_gauss", "dense", 5, 2, 0.07, 12),
        make_case(rng, "diag_dense", "diag_gauss", "dense", 4, 2, 0.1, 9),
        make_case(rng, "funnel_diag", "funnel", "diag", 6, 3, 0.05, 12),
        make_case(rng, "funnel_unit_bwd", "funnel", "unit", 4, 2, 0.03, -9),
        make_case(rng, "d128_diag", "diag_gauss", "diag", 128, 2, 0.1, 32),
    ]
    with open(os.path.join(HERE, "leapfrog_mp50.json"), "w") as f:
        json.dump(dict(generator="tests/golden/gen_closed_form.py", digits=50, cases=cases, survey_spots=spot_values()), f)
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
