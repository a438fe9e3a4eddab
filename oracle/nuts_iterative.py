"""nuts_iterative.py -- the ITERATIVE form of the NUTS transition that the CUDA kernel (K3) runs, as a readable numpy
specification for ONE chain.  *** TEST INFRASTRUCTURE ONLY *** (same rules as the rest of oracle/).

The reference builds the tree recursively (src/trajectory.jl:626-742).  The kernel cannot recurse; it walks the 2^j
leaves of a subtree in order and drives the merges like a binary counter: after leaf i, level k merges iff bit k of i
is set.  The only state kept per level k is the FIRST half-subtree waiting for its sibling ("pending[k]").  This file
states that scheme -- pending levels, "float-up" of terminated nodes, variate consumption order, the three termination
criteria and both trajectory samplers -- in ~150 lines of plain Python, so that it can be checked on the CPU against
the recursive restatements (oracle/ahmc_oracle.c and the 50-digit fixtures tests/golden/nuts_mp50.json): the algorithm
is pinned independently of any GPU run, and the tile-centric kernel planned next has a spec.

Correspondence with advancedhmc.jl_b200/csrc/ahmc_nuts_kernel.cuh: blocks (A) start a doubling, (B) one leaf,
(C) post-order merges, (D) subtree complete carry the same letters below.
"""
import math

import numpy as np


class System:
    """log-density + metric as plain numpy callables: grad(theta) -> (lp, -grad lp);  dHdr(r) -> M^-1 r."""

    def __init__(self, logp_grad, dHdr, eps, temper_alpha=0.0):
        self.logp_grad, self.dHdr, self.eps = logp_grad, dHdr, eps
        self.temper_alpha = temper_alpha  # > 0: TemperedLeapfrog(eps, alpha) (integrator.jl:174-209)

    def point(self, th, r):
        lp, g = self.logp_grad(th)
        return dict(th=th, r=r, lp=lp, g=g, lk=-0.5 * float(r @ self.dHdr(r)))

    def step(self, z, v):  # integrator.jl:216-265, n_steps = v
        e = self.eps if v > 0 else -self.eps
        sa = math.sqrt(self.temper_alpha) if self.temper_alpha > 0 else 1.0  # a 1-step `step`: multiply before, divide after
        r = z["r"] * sa - e / 2 * z["g"]
        th = z["th"] + e * self.dHdr(r)
        lp, g = self.logp_grad(th)
        r = (r - e / 2 * g) / sa
        return dict(th=th, r=r, lp=lp, g=g, lk=-0.5 * float(r @ self.dHdr(r)))


def logaddexp(a, b):
    m = max(a, b)
    return m + math.log1p(math.exp(-abs(a - b)))


def maxabs(a, b):
    return a if abs(a) > abs(b) else b


def transition(S, z0, dirs, variates, sampler="multinomial", criterion="generalised", max_depth=10, delta_max=1000.0,
               linear_accept=False, max_weights=False):
    """-> (zcand, stats dict, number of variates used).  dirs: direction bits (1 = left), variates: the tape.
    linear_accept: take the multinomial decisions in the probability domain -- `u < w_p / (w_p + w_c)` with
    u = exp(-randexp) -- instead of the reference's `lw < lw_p + randexp` (mathematically the same event; it would save
    the kernel one log per combine because exp(-|lw_p - lw_c|) is already computed by logaddexp and the Philox stream
    yields u directly).  tests/test_oracle.py checks that both forms decide identically on every test tree.
    max_weights (implies the probability-domain decisions): carry every multinomial weight as (m, w) with
    log-weight = m + log(w), m = the largest leaf log-weight under the node and w in [1, #leaves]; combining two nodes
    costs one exp and no log at all -- the form staged in the kernel behind AHMC_NUTS_FASTDRAW."""
    nvar = ndir = 0

    def draw():
        nonlocal nvar
        nvar += 1
        return variates[nvar - 1]

    H0 = -(z0["lp"] + z0["lk"])
    LEFT, RIGHT, rho_tree = z0, z0, z0["r"].copy()
    zc = z0
    if sampler == "slice":
        lu, lw_tree = (z0["lp"] + z0["lk"]) - draw(), 1.0
    else:
        lu, lw_tree = None, 0.0
    sa_tree = dh_tree = 0.0
    ww_tree = 1.0
    na_tree = j = 0
    term_dyn = term_num = False
    pending = {}  # level k -> dict(rho, rfirst, rlast, thfirst, cand, lw, sa, na, dh)
    while not (term_dyn or term_num) and j < max_depth:
        # ---------------------------------------------------------------- (A) start a doubling
        v = -1 if dirs[ndir] else 1
        ndir += 1
        s = LEFT if v < 0 else RIGHT
        jsub, i = j, 0
        while True:
            # ------------------------------------------------------------ (B) one leaf
            s = S.step(s, v)
            nE = s["lp"] + s["lk"]
            H1 = -nE
            dH = H1 - H0
            if sampler == "slice":
                lw_c, tnum_c = (1.0 if lu <= nE else 0.0), not (lu < delta_max + -H1)
            else:
                lw_c, tnum_c = H0 + nE, not (-H0 < delta_max + -H1)
            sa_c, na_c, dh_c, tdyn_c = math.exp(min(0.0, -dH)), 1, dH, False
            ww_c = 1.0  # max_weights: (lw_c, ww_c) = (m, w)
            node = dict(rho=s["r"].copy(), rfirst=s["r"], rlast=s["r"], thfirst=s["th"], cand=s)
            # ------------------------------------------------------------ (C) post-order merges, binary-counter style
            k, complete = 0, False
            while True:
                if k == jsub:
                    complete = True
                    break
                bit = (i >> k) & 1
                if bit:  # a first half is waiting at level k: combine (first = pending, second = current)
                    F = pending.pop(k)
                    if criterion == "strict":  # :579-613, direction independent in (first, second) form
                        ra, rb = F["rho"] + node["rfirst"], node["rho"] + F["rlast"]
                        extra = (ra @ S.dHdr(F["rfirst"]) <= 0 or ra @ S.dHdr(node["rfirst"]) <= 0
                                 or rb @ S.dHdr(s["r"]) <= 0 or rb @ S.dHdr(F["rlast"]) <= 0)
                    else:
                        extra = False
                    rho = node["rho"] + F["rho"]
                    if criterion == "classic":  # :551-557; q = theta_left - theta_right
                        q = (F["thfirst"] - s["th"]) if v > 0 else (s["th"] - F["thfirst"])
                        uturn = q @ S.dHdr(F["rfirst"]) >= 0 or q @ S.dHdr(s["r"]) >= 0
                    else:
                        uturn = rho @ S.dHdr(F["rfirst"]) <= 0 or rho @ S.dHdr(s["r"]) <= 0 or extra
                    u = draw()
                    if sampler == "slice":  # :178-183
                        n = F["lw"] + lw_c
                        cand = F["cand"] if n * u < F["lw"] else node["cand"]
                        lw_c = n
                    elif max_weights:  # :191-195 on (m, w) pairs: one exp, no log
                        d = F["lw"] - lw_c
                        t = math.exp(-abs(d)) if d == d else float("nan")
                        if d >= 0:
                            w_new, p_first, m_new = F["ww"] + ww_c * t, None, F["lw"]
                            p_first = F["ww"] / w_new
                        else:
                            w_new, m_new = F["ww"] * t + ww_c, lw_c
                            p_first = F["ww"] * t / w_new
                        take = (d == d) and (math.exp(-u) < p_first)
                        cand = F["cand"] if take else node["cand"]
                        lw_c, ww_c = m_new, w_new
                    else:  # :191-195
                        lw = logaddexp(F["lw"], lw_c)
                        if linear_accept:
                            t = math.exp(-abs(F["lw"] - lw_c))
                            p_first = 1.0 / (1.0 + t) if F["lw"] >= lw_c else t / (1.0 + t)
                            take = math.exp(-u) < p_first
                        else:
                            take = lw < F["lw"] + u
                        cand = F["cand"] if take else node["cand"]
                        lw_c = lw
                    sa_c = F["sa"] + sa_c if v > 0 else sa_c + F["sa"]
                    na_c += F["na"]
                    dh_c = maxabs(F["dh"], dh_c) if v > 0 else maxabs(dh_c, F["dh"])
                    tdyn_c = tdyn_c or uturn
                    node = dict(rho=rho, rfirst=F["rfirst"], rlast=s["r"], thfirst=F["thfirst"], cand=cand)
                    k += 1
                elif tnum_c or tdyn_c:  # a terminated first half is returned as is (:652): it "floats" up a level
                    k += 1
                else:  # first half of level k: park it and go build its sibling
                    pending[k] = dict(node, lw=lw_c, ww=ww_c, sa=sa_c, na=na_c, dh=dh_c)
                    break
            if complete:
                break
            i += 1
        # ---------------------------------------------------------------- (D) subtree complete
        sub_term = tnum_c or tdyn_c
        if not sub_term:
            j += 1
            u = draw()
            if sampler == "slice":
                accept = lw_tree * u < lw_c
            elif max_weights:  # lw_tree < lw_c + randexp  <=>  u < (w_c / w_T) exp(m_c - m_T)
                accept = math.exp(-u) < (ww_c / ww_tree) * math.exp(lw_c - lw_tree)
            elif linear_accept:
                accept = math.exp(-u) < math.exp(min(0.0, lw_c - lw_tree))
            else:
                accept = lw_tree < lw_c + u
            if accept:
                zc = node["cand"]
        near, far = (LEFT, RIGHT) if v < 0 else (RIGHT, LEFT)
        if criterion == "strict":
            rx, ry = rho_tree + node["rfirst"], near["r"] + node["rho"]
            extra = (rx @ S.dHdr(far["r"]) <= 0 or rx @ S.dHdr(node["rfirst"]) <= 0
                     or ry @ S.dHdr(near["r"]) <= 0 or ry @ S.dHdr(s["r"]) <= 0)
        else:
            extra = False
        rho_tree = rho_tree + node["rho"]
        if v < 0:
            LEFT = s
        else:
            RIGHT = s
        if criterion == "classic":
            q = (far["th"] - s["th"]) if v > 0 else (s["th"] - far["th"])
            uturn = q @ S.dHdr(far["r"]) >= 0 or q @ S.dHdr(s["r"]) >= 0
        else:
            uturn = rho_tree @ S.dHdr(far["r"]) <= 0 or rho_tree @ S.dHdr(s["r"]) <= 0 or extra
        sa_tree = sa_c + sa_tree if v < 0 else sa_tree + sa_c
        na_tree += na_c
        dh_tree = maxabs(dh_c, dh_tree) if v < 0 else maxabs(dh_tree, dh_c)
        if sampler == "slice":
            lw_tree = lw_tree + lw_c
        elif max_weights:
            d = lw_tree - lw_c
            t = math.exp(-abs(d)) if d == d else float("nan")
            if d >= 0:
                ww_tree = ww_tree + ww_c * t
            else:
                lw_tree, ww_tree = lw_c, ww_tree * t + ww_c
        else:
            lw_tree = logaddexp(lw_tree, lw_c)
        term_dyn = term_dyn or tdyn_c or uturn
        term_num = term_num or tnum_c
        pending.clear()
    H = -(zc["lp"] + zc["lk"])
    return zc, dict(n_steps=na_tree, tree_depth=j, numerical_error=term_num, acceptance_rate=sa_tree / na_tree,
                    hamiltonian_energy_error=H - H0, max_hamiltonian_energy_error=dh_tree), nvar
