"""gen_nuts_mp.py -- known-answer fixtures for NUTS transitions: tests/golden/nuts_mp50.json.

A SECOND, independent restatement of the reference's dynamic-trajectory transition, written directly from the Julia
source in its own recursive shape (`build_tree` / `transition`, src/trajectory.jl:626-742) and evaluated in 50-digit
mpmath arithmetic.  It shares no code with oracle/ahmc_oracle.c (C, fp64, written separately) nor with the CUDA kernel
(iterative, binary-counter merges), so agreement of all three on the discrete outcome of every comparison (tree depth,
number of leapfrog steps, selected candidate, divergence flag) and on the continuous outputs to 1e-10 pins the tree
logic about as well as is possible without running Julia (the reference holds no golden vectors for this path,
SURVEY.md section 8c).  Covered: MultinomialTS / SliceTS (:102-206), GeneralisedNoUTurn / ClassicNoUTurn /
StrictGeneralisedNoUTurn (:551-617), numerical termination (:500-507), all four built-in targets, Unit / Diag / Dense
metrics.  Randomness comes from tapes: one direction bit per doubling (`rand(rng, Bool)`, :693) and one variate per
`combine(rng, ...)` / `mh_accept` in the reference's consumption order (randexp for MultinomialTS; for SliceTS one
randexp for the slice variable, then rand() uniforms).  The momentum is given (no refresh), so the fixture isolates
the tree.

Every comparison the algorithm makes is recorded with its margin; a case whose smallest relative margin is below 1e-7
is rejected at generation time, so fp64 implementations cannot legitimately take a different branch.

Run:  python tests/golden/gen_nuts_mp.py      (rewrites nuts_mp50.json deterministically)
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


class Z:
    """PhasePoint (src/hamiltonian.jl:88-107): theta, r, lp, g = -grad lp, lk."""

    __slots__ = ("th", "r", "lp", "g", "lk")

    def __init__(self, th, r, lp, g, lk):
        self.th, self.r, self.lp, self.g, self.lk = th, r, lp, g, lk


class Ctx:
    def __init__(self, kind, D, p0, p1, c0, mkind, Minv, eps, sampler, criterion, max_depth, delta_max, dirs, variates):
        self.kind, self.D, self.p0, self.p1, self.c0 = kind, D, p0, p1, c0
        self.mkind, self.Minv, self.eps = mkind, Minv, eps
        self.sampler, self.criterion, self.max_depth, self.delta_max = sampler, criterion, max_depth, delta_max
        self.dirs, self.variates = list(dirs), list(variates)
        self.temper = None  # mp.mpf alpha: the integrator is TemperedLeapfrog(eps, alpha) (integrator.jl:174-209)
        self.n_dir = self.n_var = 0
        self.min_margin = mp.mpf(1)

    # ---- randomness, in consumption order
    def rand_bool(self):
        b = self.dirs[self.n_dir]
        self.n_dir += 1
        return bool(b)

    def variate(self):
        v = self.variates[self.n_var]
        self.n_var += 1
        return v

    # ---- a comparison whose outcome fp64 must reproduce
    def less(self, a, b, strict=True):
        scale = max(abs(a), abs(b), mp.mpf(1))
        self.min_margin = min(self.min_margin, abs(a - b) / scale)
        return a < b if strict else a <= b

    def dHdr(self, r):
        return dHdr_mp(self.mkind, self.Minv, r, self.D)

    def phasepoint(self, th, r):
        lp, grad = logp_grad_mp(self.kind, self.D, self.p0, self.p1, self.c0, th)
        dr = self.dHdr(r)
        lk = -sum(r[d] * dr[d] for d in range(self.D)) / 2
        return Z(th, r, lp, [-x for x in grad], lk)

    def step(self, z, v):  # integrator.jl:216-265 with n_steps = v = +-1
        e = self.eps if v > 0 else -self.eps
        D = self.D
        r0 = z.r
        if self.temper is not None:  # n_steps = 1, i = 1: i_temper = 1 <= 1 -> multiply before the first half kick ...
            r0 = [x * mp.sqrt(self.temper) for x in r0]
        r = [r0[d] - e / 2 * z.g[d] for d in range(D)]
        dr = self.dHdr(r)
        th = [z.th[d] + e * dr[d] for d in range(D)]
        lp, grad = logp_grad_mp(self.kind, D, self.p0, self.p1, self.c0, th)
        g = [-x for x in grad]
        r = [r[d] - e / 2 * g[d] for d in range(D)]
        if self.temper is not None:  # ... i_temper = 2 > 1 -> divide after the second
            r = [x / mp.sqrt(self.temper) for x in r]
        dr = self.dHdr(r)
        return Z(th, r, lp, g, -sum(r[d] * dr[d] for d in range(D)) / 2)


def neg_energy(z):
    return z.lp + z.lk


def dot(a, b):
    return sum(x * y for x, y in zip(a, b))


def logaddexp(a, b):
    m = max(a, b)
    return m + mp.log1p(mp.e ** (-abs(a - b)))


def maxabs(a, b):  # :526
    return a if abs(a) > abs(b) else b


class Tree:  # BinaryTree (:512-519)
    __slots__ = ("zleft", "zright", "rho", "sum_a", "n_a", "dH_max")

    def __init__(self, zl, zr, rho, sa, na, dh):
        self.zleft, self.zright, self.rho, self.sum_a, self.n_a, self.dH_max = zl, zr, rho, sa, na, dh


def combine_tree(tl, tr):  # :533-542
    return Tree(tl.zleft, tr.zright, [a + b for a, b in zip(tl.rho, tr.rho)], tl.sum_a + tr.sum_a, tl.n_a + tr.n_a,
                maxabs(tl.dH_max, tr.dH_max))


class Sampler:  # MultinomialTS: w = log weight; SliceTS: lu, w = n
    __slots__ = ("zcand", "w", "lu")

    def __init__(self, zcand, w, lu=None):
        self.zcand, self.w, self.lu = zcand, w, lu


def leaf_sampler(c, s, H0, z):
    if c.sampler == "slice":  # :164-166
        return Sampler(z, 1 if c.less(s.lu, neg_energy(z), strict=False) else 0, s.lu)
    return Sampler(z, H0 + neg_energy(z))  # :174-176


def combine_rng(c, s1, s2):
    if c.sampler == "slice":  # :178-183
        n = s1.w + s2.w
        u = c.variate()
        return Sampler(s1.zcand if c.less(n * u, mp.mpf(s1.w)) else s2.zcand, n, s1.lu)
    lw = logaddexp(s1.w, s2.w)  # :191-195
    ex = c.variate()
    return Sampler(s1.zcand if c.less(lw, s1.w + ex) else s2.zcand, lw)


def combine_cand(c, zcand, s1, s2):
    if c.sampler == "slice":  # :185-189
        return Sampler(zcand, s1.w + s2.w, s1.lu)
    return Sampler(zcand, logaddexp(s1.w, s2.w))  # :197-200


def mh_accept(c, s, s2):
    if c.sampler == "slice":  # :202
        return c.less(s.w * c.variate(), mp.mpf(s2.w))
    return c.less(s.w, s2.w + c.variate())  # :204-206


def termination_numerical(c, s, H0, H1):
    if c.sampler == "slice":  # :500-502
        return not c.less(s.lu, c.delta_max + -H1)
    return not c.less(-H0, c.delta_max + -H1)  # :503-507


def gen_uturn(c, rho, p_minus, p_plus):  # :615-617
    a, b = dot(rho, p_minus), dot(rho, p_plus)
    ta = c.less(a, mp.mpf(0), strict=False)
    tb = c.less(b, mp.mpf(0), strict=False)
    return ta or tb


def isterminated(c, t, tleft, tright):
    if c.criterion == "classic":  # :551-557
        dth = [a - b for a, b in zip(t.zright.th, t.zleft.th)]
        s1 = dot(dth, c.dHdr([-x for x in t.zleft.r]))
        s2 = dot([-x for x in dth], c.dHdr(t.zright.r))
        a = c.less(mp.mpf(0), s1, strict=False)  # s1 >= 0
        b = c.less(mp.mpf(0), s2, strict=False)
        return a or b
    s = gen_uturn(c, t.rho, c.dHdr(t.zleft.r), c.dHdr(t.zright.r))  # :566-570
    if c.criterion == "strict":  # :579-613
        rho = [a + b for a, b in zip(tleft.rho, tright.zleft.r)]
        s2 = gen_uturn(c, rho, c.dHdr(t.zleft.r), c.dHdr(tright.zleft.r))
        rho = [a + b for a, b in zip(tleft.zright.r, tright.rho)]
        s3 = gen_uturn(c, rho, c.dHdr(tleft.zright.r), c.dHdr(t.zright.r))
        s = s or s2 or s3
    return s


def build_tree(c, z, sampler, v, j, H0):  # :626-675; returns (tree, sampler, (dynamic, numerical))
    if j == 0:
        z1 = c.step(z, v)
        H1 = -neg_energy(z1)
        dH = H1 - H0
        alpha = mp.e ** min(mp.mpf(0), -dH)
        s1 = leaf_sampler(c, sampler, H0, z1)
        return Tree(z1, z1, list(z1.r), alpha, 1, dH), s1, (False, termination_numerical(c, s1, H0, H1))
    tree1, s1, term1 = build_tree(c, z, sampler, v, j - 1, H0)
    if not (term1[0] or term1[1]):
        if v == -1:
            tree2, s2, term2 = build_tree(c, tree1.zleft, sampler, v, j - 1, H0)
            tl, tr = tree2, tree1
        else:
            tree2, s2, term2 = build_tree(c, tree1.zright, sampler, v, j - 1, H0)
            tl, tr = tree1, tree2
        tree1 = combine_tree(tl, tr)
        s1 = combine_rng(c, s1, s2)
        dyn = isterminated(c, tree1, tl, tr)
        term1 = (term1[0] or term2[0] or dyn, term1[1] or term2[1])
    return tree1, s1, term1


def transition(c, z0):  # :677-742
    H0 = -neg_energy(z0)
    tree = Tree(z0, z0, list(z0.r), mp.mpf(0), 0, mp.mpf(0))
    if c.sampler == "slice":
        sampler = Sampler(z0, 1, neg_energy(z0) - c.variate())  # :144-145
    else:
        sampler = Sampler(z0, mp.mpf(0))  # :155
    term = (False, False)
    zcand = z0
    j = 0
    while not (term[0] or term[1]) and j < c.max_depth:
        if c.rand_bool():
            t1, s1, tm1 = build_tree(c, tree.zleft, sampler, -1, j, H0)
            tl, tr = t1, tree
        else:
            t1, s1, tm1 = build_tree(c, tree.zright, sampler, 1, j, H0)
            tl, tr = tree, t1
        if not (tm1[0] or tm1[1]):
            j += 1
            if mh_accept(c, sampler, s1):
                zcand = s1.zcand
        tree = combine_tree(tl, tr)
        sampler = combine_cand(c, zcand, sampler, s1)
        dyn = isterminated(c, tree, tl, tr)
        term = (term[0] or tm1[0] or dyn, term[1] or tm1[1])
    H = -neg_energy(zcand)
    return zcand, dict(n_steps=tree.n_a, acceptance_rate=tree.sum_a / tree.n_a, log_density=zcand.lp,
                       hamiltonian_energy=H, hamiltonian_energy_error=H - H0, max_hamiltonian_energy_error=tree.dH_max,
                       tree_depth=j, numerical_error=bool(term[1]))


def make_case(rng, name, kind, mkind, D, N, eps, sampler, criterion, max_depth=6, delta_max=1000.0, scale=1.0, temper=None):
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
    n_var = 1 << max_depth
    while True:  # redraw until every comparison has a comfortable margin
        theta = rng.normal(size=(N, D)) * scale
        r = rng.normal(size=(N, D))
        dirs = rng.integers(0, 2, size=(N, max_depth + 1)).astype(np.uint8)
        var = rng.exponential(size=(N, n_var))
        if sampler == "slice":
            var[:, 1:] = rng.uniform(size=(N, n_var - 1))
        out = dict(theta=[], r=[], lp_gradient=[], lp_value=[], lk_value=[], n_steps=[], tree_depth=[], numerical_error=[],
                   acceptance_rate=[], hamiltonian_energy_error=[], max_hamiltonian_energy_error=[], variates_used=[])
        margin = mp.mpf(1)
        for ch in range(N):
            c = Ctx(kind, D, pp0, pp1, mp.mpf(c0), mkind, Mm, mp.mpf(float(eps)), sampler, criterion, max_depth,
                    mp.mpf(float(delta_max)), dirs[ch], mpf_list(var[ch]))
            c.temper = None if temper is None else mp.mpf(float(temper))
            z0 = c.phasepoint(mpf_list(theta[ch]), mpf_list(r[ch]))
            zc, st = transition(c, z0)
            margin = min(margin, c.min_margin)
            out["theta"].append([float(x) for x in zc.th])
            out["r"].append([float(x) for x in zc.r])
            out["lp_gradient"].append([float(x) for x in zc.g])
            out["lp_value"].append(float(zc.lp))
            out["lk_value"].append(float(zc.lk))
            out["variates_used"].append(c.n_var)
            for k in ("n_steps", "tree_depth"):
                out[k].append(int(st[k]))
            out["numerical_error"].append(bool(st["numerical_error"]))
            for k in ("acceptance_rate", "hamiltonian_energy_error", "max_hamiltonian_energy_error"):
                out[k].append(float(st[k]))
        if margin > mp.mpf("1e-7"):
            break
    tolist = lambda a: None if a is None else np.asarray(a).tolist()
    return dict(name=name, model=kind, metric=mkind, D=D, N=N, eps=float(eps), sampler=sampler, criterion=criterion,
                max_depth=max_depth, delta_max=float(delta_max), temper_alpha=0.0 if temper is None else float(temper), p0=tolist(p0), p1=tolist(p1), c0=c0, Minv=tolist(Minv),
                theta0=theta.tolist(), r0=r.tolist(), dirs=dirs.tolist(), variates=var.tolist(),
                min_margin=float(margin), expect=out)


def main():
    rng = np.random.Generator(np.random.PCG64(20260924))
    cases = [
        make_case(rng, "mn_gen_stdnormal_unit", "std_normal", "unit", 4, 6, 0.35, "multinomial", "generalised"),
        make_case(rng, "mn_gen_diag_diag", "diag_gauss", "diag", 5, 6, 0.3, "multinomial", "generalised"),
        make_case(rng, "mn_gen_funnel_diag", "funnel", "diag", 4, 6, 0.25, "multinomial", "generalised", scale=0.7),
        make_case(rng, "mn_gen_dense_dense", "dense_gauss", "dense", 4, 5, 0.3, "multinomial", "generalised"),
        make_case(rng, "mn_gen_maxdepth", "std_normal", "unit", 3, 4, 0.02, "multinomial", "generalised", max_depth=4),
        make_case(rng, "mn_gen_divergent", "funnel", "unit", 3, 6, 1.6, "multinomial", "generalised", delta_max=3.0, scale=1.5),
        make_case(rng, "slice_gen_diag_diag", "diag_gauss", "diag", 5, 6, 0.3, "slice", "generalised"),
        make_case(rng, "slice_gen_divergent", "funnel", "unit", 3, 6, 1.6, "slice", "generalised", delta_max=3.0, scale=1.5),
        make_case(rng, "mn_classic_diag_diag", "diag_gauss", "diag", 5, 6, 0.3, "multinomial", "classic"),
        make_case(rng, "mn_classic_dense_dense", "dense_gauss", "dense", 4, 5, 0.3, "multinomial", "classic"),
        make_case(rng, "mn_strict_diag_diag", "diag_gauss", "diag", 5, 6, 0.3, "multinomial", "strict"),
        make_case(rng, "mn_strict_funnel_unit", "funnel", "unit", 4, 6, 0.3, "multinomial", "strict", scale=0.7),
        make_case(rng, "slice_strict_stdnormal_diag", "std_normal", "diag", 4, 6, 0.35, "slice", "strict"),
        make_case(rng, "slice_classic_diag_unit", "diag_gauss", "unit", 4, 6, 0.3, "slice", "classic"),
        # TemperedLeapfrog as the integrator: every leaf is a 1-step `step` (appended: the cases above keep their streams)
        make_case(rng, "mn_gen_diag_diag_tempered", "diag_gauss", "diag", 5, 6, 0.3, "multinomial", "generalised", temper=1.06),
        make_case(rng, "mn_gen_dense_dense_tempered", "dense_gauss", "dense", 4, 5, 0.3, "multinomial", "generalised", temper=0.95),
    ]
    with open(os.path.join(HERE, "nuts_mp50.json"), "w") as f:
        json.dump(dict(generator="tests/golden/gen_nuts_mp.py", digits=50, cases=cases), f)
    for c in cases:
        e = c["expect"]
        print(f"{c['name']:32s} depth {e['tree_depth']} steps {e['n_steps']} diverged {sum(e['numerical_error'])} "
              f"min margin {c['min_margin']:.2e}")


if __name__ == "__main__":
    main()
