"""CPU tests: pin the oracle (C restatement + numpy twin) against the 50-digit known answers and the
identities the reference's own tests assert.  No GPU, no reference at run time."""
import math

import numpy as np
import pytest

from oracle import oracle_c as oc
from oracle import oracle_np as onp
from tests.helpers import METRIC_KINDS, MODEL_KINDS, case_arrays, golden_cases, rel_err

GOLD = golden_cases()


def _c_objs(case, a):
    model = oc.Model(MODEL_KINDS[case["model"]], case["D"], a["p0"], a["p1"], case["c0"])
    metric = oc.Metric(METRIC_KINDS[case["metric"]], a["Minv"])
    return model, metric


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_c_oracle_matches_mp50(case):
    a = case_arrays(case)
    model, metric = _c_objs(case, a)
    z0 = oc.phasepoint(model, metric, a["theta0"], a["r0"])
    z1, status, done = oc.leapfrog(model, metric, a["eps"], z0, case["n_steps"], case["temper_alpha"] or 0.0)
    assert (status == 0).all() and (done == abs(case["n_steps"])).all()
    tol = 2e-12  # fp64 op-order result vs exact: 1e-16..1e-13 (chaotic funnel amplifies a little)
    assert rel_err(z1.theta, a["theta"]) < tol
    assert rel_err(z1.r, a["r"]) < tol
    assert rel_err(z1.lp_gradient, a["lp_gradient"]) < tol
    assert rel_err(z1.lp_value, a["lp_value"]) < tol
    assert rel_err(z1.lk_value, a["lk_value"]) < tol


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_numpy_twin_matches_c_oracle(case):
    a = case_arrays(case)
    model, metric = _c_objs(case, a)
    z0 = oc.phasepoint(model, metric, a["theta0"], a["r0"])
    z1, _, _ = oc.leapfrog(model, metric, a["eps"], z0, case["n_steps"], case["temper_alpha"] or 0.0)
    nm = onp.Model(MODEL_KINDS[case["model"]], case["D"], a["p0"], a["p1"], case["c0"])
    me = onp.Metric(METRIC_KINDS[case["metric"]], a["Minv"])
    y0 = onp.phasepoint(nm, me, a["theta0"], a["r0"])
    assert rel_err(y0.lp_value, z0.lp_value) < 1e-14 and rel_err(y0.lk_value, z0.lk_value) < 1e-14
    y1 = onp.step(nm, me, a["eps"], y0, case["n_steps"], temper_alpha=case["temper_alpha"])
    for f in ("theta", "r", "lp_gradient", "lp_value", "lk_value", "lk_gradient"):
        assert rel_err(getattr(y1, f), getattr(z1, f)) < 5e-13, f


def test_survey_spot_values():
    """SURVEY 8c spot vectors: fp64 op-order results recorded there to 16 digits."""
    model = oc.Model(oc.DIAG_GAUSS, 1, [0.0], [1.0])
    metric = oc.Metric(oc.DIAG, [1.0])
    z0 = oc.phasepoint(model, metric, np.array([[1.0]]), np.array([[0.5]]))
    z1, _, _ = oc.leapfrog(model, metric, 0.1, z0, 32)
    assert z1.theta[0, 0] == pytest.approx(-1.0281066785335988139, rel=1e-14)
    assert z1.r[0, 0] == pytest.approx(-0.43947601289572651662, rel=1e-14)
    assert z1.lp_value[0] == pytest.approx(-0.5 * z1.theta[0, 0] ** 2, rel=1e-15)
    assert z1.lk_value[0] == pytest.approx(-0.09656958295536232, rel=1e-13)
    s = GOLD["survey_spots"]
    assert s[0]["theta"].startswith("-1.02810667853359881") and s[2]["r"].startswith("1.30542169799302271")


def test_energy_identities():
    """test/hamiltonian.jl:54-79: neg_energy / dHdr identities for Unit, Diag, Dense."""
    rng = np.random.default_rng(1)
    D = 5
    for _ in range(10):
        r = rng.normal(size=D)
        mu = oc.Metric(oc.UNIT)
        assert -oc.neg_kinetic(mu, r) == np.sum(r * r) / 2
        assert (oc.dHdr(mu, r) == r).all()
        Mi = 1 + np.abs(rng.normal(size=D))
        me = oc.Metric(oc.DIAG, Mi)
        assert (oc.dHdr(me, r) == Mi * r).all()
        assert -oc.neg_kinetic(me, r) == pytest.approx(r @ np.diag(Mi) @ r / 2, rel=1e-14)
        m = rng.normal(size=(D, D))
        Md = m.T @ m
        md = oc.Metric(oc.DENSE, Md)
        assert np.allclose(oc.dHdr(md, r), Md @ r, rtol=1e-13)
        assert -oc.neg_kinetic(md, r) == pytest.approx(r @ Md @ r / 2, rel=1e-13)


def test_step_n_equals_n_steps_of_one():
    """test/integrator.jl:17-32 (there atol 5e-3; here bit-identical by construction)."""
    rng = np.random.default_rng(2)
    D, N = 5, 7
    model, metric = oc.Model(oc.STD_NORMAL, D), oc.Metric(oc.UNIT)
    z = oc.phasepoint(model, metric, rng.normal(size=(D, N)), rng.normal(size=(D, N)))
    zn, _, _ = oc.leapfrog(model, metric, 0.01, z, 10)
    zl = z
    for _ in range(10):
        zl, _, _ = oc.leapfrog(model, metric, 0.01, zl, 1)
    assert (zl.theta == zn.theta).all() and (zl.r == zn.r).all()


def test_harmonic_oscillator_invariants():
    """test/integrator.jl:108-153: radius and energy constant within 2e-3 over 10^4 steps, eps=0.01."""
    model, metric = oc.Model(oc.STD_NORMAL, 1), oc.Metric(oc.UNIT)
    rng = np.random.default_rng(3)
    z = oc.phasepoint(model, metric, rng.normal(size=(1, 1)), rng.normal(size=(1, 1)))
    traj, done = oc.leapfrog_trajectory(model, metric, 0.01, z, 10_000)
    assert done[0] == 10_000
    q, p = traj["theta"][0, 0, 999:], traj["r"][0, 0, 999:]
    H = -(traj["lp_value"][0, 999:] + traj["lk_value"][0, 999:])
    rs = np.sqrt(q**2 + p**2)
    assert np.all(np.abs(rs - rs.mean()) < 2e-3) and np.all(np.abs(H - H.mean()) < 2e-3)


def test_reversibility_backward_steps():
    """n_steps<0 integrates backward (integrator.jl:221-226): fwd n then bwd n returns to start."""
    rng = np.random.default_rng(4)
    D, N = 6, 3
    s = np.exp(rng.uniform(-1, 1, D))
    model, metric = oc.Model(oc.DIAG_GAUSS, D, rng.normal(size=D), s), oc.Metric(oc.DIAG, s * s)
    z = oc.phasepoint(model, metric, rng.normal(size=(D, N)), rng.normal(size=(D, N)))
    z1, _, _ = oc.leapfrog(model, metric, 0.1, z, 17)
    z2, _, _ = oc.leapfrog(model, metric, 0.1, z1, -17)
    assert rel_err(z2.theta, z.theta) < 1e-13 and rel_err(z2.r, z.r) < 1e-13


def test_nonfinite_break_per_chain_and_compat():
    """integrator.jl:252-258 + hamiltonian.jl:95-104,141-142 (quirk Q1)."""
    D, N = 3, 4
    model, metric = oc.Model(oc.STD_NORMAL, D), oc.Metric(oc.UNIT)
    th = np.ones((D, N))
    th[:, 2] = 1e200  # kinetic/potential overflow for chain 2 at the first step
    z = oc.phasepoint(model, metric, th, np.ones((D, N)))
    z1, status, done = oc.leapfrog(model, metric, 0.1, z, 5)
    assert list(status) == [0, 0, 1, 0] and list(done) == [5, 5, 1, 5]
    assert z1.lp_value[2] == -np.inf
    z2, status2, done2 = oc.leapfrog(model, metric, 0.1, z, 5, compat_break_all=True)
    assert list(done2) == [1, 1, 1, 1] and list(status2) == [0, 0, 1, 0]
    ref, _, _ = oc.leapfrog(model, metric, 0.1, z, 1)
    assert (z2.theta == ref.theta).all()
    # numpy twin = the reference's matrix mode: breaks all chains too
    y = onp.step(onp.Model(onp.STD_NORMAL, D), onp.Metric(onp.UNIT), 0.1,
                 onp.phasepoint(onp.Model(onp.STD_NORMAL, D), onp.Metric(onp.UNIT), th, np.ones((D, N))), 5)
    assert np.allclose(y.theta[:, 0], z2.theta[:, 0], rtol=1e-15)


def test_chain_independence():
    """test/sampler-vec.jl:69-80, test/metric.jl:4-22: identical columns stay identical; a chain's
    result does not depend on its neighbours."""
    rng = np.random.default_rng(5)
    D = 5
    th, r = rng.normal(size=(D, 1)), rng.normal(size=(D, 1))
    model, metric = oc.Model(oc.FUNNEL, D), oc.Metric(oc.UNIT)
    z = oc.phasepoint(model, metric, np.repeat(th, 6, 1), np.repeat(r, 6, 1))
    z1, _, _ = oc.leapfrog(model, metric, 0.05, z, 11)
    assert all((z1.theta[:, j] == z1.theta[:, 0]).all() for j in range(6))
    zs = oc.phasepoint(model, metric, th, r)
    z1s, _, _ = oc.leapfrog(model, metric, 0.05, zs, 11)
    assert (z1s.theta[:, 0] == z1.theta[:, 3]).all()


def test_hmc_transition_c_vs_numpy():
    rng = np.random.default_rng(6)
    D, N = 6, 9
    s = np.exp(rng.uniform(-0.5, 0.5, D))
    a = (oc.DIAG_GAUSS, D, rng.normal(size=D), s)
    model, metric = oc.Model(*a), oc.Metric(oc.DIAG, s * s)
    nm, me = onp.Model(*a), onp.Metric(onp.DIAG, s * s)
    th = rng.normal(size=(D, N))
    z = oc.phasepoint(model, metric, th, rng.normal(size=(D, N)))
    nt, et = rng.normal(size=(D, N)), rng.exponential(size=N) * 0.05
    zc, st = oc.hmc_transition(model, metric, 0.45, 12, z, nt, et)
    yn, sn = onp.hmc_transition(nm, me, 0.45, 12, onp.phasepoint(nm, me, np.asfortranarray(th), z.r), np.asfortranarray(nt), et)
    assert 0 < st.is_accept.sum() < N  # both branches exercised
    assert (st.is_accept.astype(bool) == sn["is_accept"]).all()
    for f in ("theta", "r", "lp_gradient", "lp_value", "lk_value"):
        assert rel_err(getattr(zc, f), getattr(yn, f)) < 1e-13, f
    assert rel_err(st.acceptance_rate, sn["acceptance_rate"]) < 1e-12
    assert rel_err(st.hamiltonian_energy_error, sn["hamiltonian_energy_error"]) < 1e-9
    # rejected chains keep theta, momentum flipped (trajectory.jl:283,312-332)
    rej = ~st.is_accept.astype(bool)
    assert (zc.theta[:, rej] == th[:, rej]).all()


def test_stan_window_schedule_pin():
    """test/adaptation.jl:131-151: n_adapts=1000 -> start 76, end 950, splits [100,150,250,450,950]."""
    assert oc.stan_windows(1000) == (76, 950, [100, 150, 250, 450, 950])
    ws, we, sp = oc.stan_windows(100)
    assert ws == 76 and we == 50 and sp == []


def test_welford_against_naive():
    """test/adaptation.jl:63-99: Welford var/cov equal the naive estimators (+ Stan regularisation)."""
    rng = np.random.default_rng(7)
    D, n = 4, 200
    X = rng.normal(size=(n, D)) * np.array([1.0, 2.0, 0.5, 3.0])
    wv, wc = oc.WelfordVar((D,)), oc.WelfordCov(D)
    for x in X:
        wv.push(x)
        wc.push(x)
    reg = lambda M: n / ((n + 5.0)) * M + 1e-3 * (5.0 / (n + 5.0))
    assert np.allclose(wv.estimate(), reg(X.var(axis=0, ddof=1)), rtol=1e-12)
    C = np.cov(X.T, ddof=1)
    assert np.allclose(wc.estimate(), n / (n + 5.0) * C + 1e-3 * (5.0 / (n + 5.0)) * np.eye(D), rtol=1e-11, atol=1e-14)


def test_dual_averaging_closed_form():
    """stepsize.jl:178-210 restated inline in numpy and compared; finalize -> exp(x_bar)."""
    da = oc.DualAveraging([0.1, 0.3], delta=0.8)
    mu = np.log(10 * np.array([0.1, 0.3]))
    xb, Hb = np.zeros(2), np.zeros(2)
    rng = np.random.default_rng(8)
    for m in range(1, 40):
        alpha = rng.uniform(0, 1.3, size=2)
        da.adapt(alpha)
        eta = 1.0 / (m + 10.0)
        Hb = (1 - eta) * Hb + eta * (0.8 - np.minimum(1.0, alpha))
        x = mu - Hb * (np.sqrt(m) / 0.05)
        ex = m ** (-0.75)
        xb = (1 - ex) * xb + ex * x
        assert np.allclose(da.eps, np.exp(x), rtol=1e-14) and da.m == m
    da.finalize()
    assert np.allclose(da.eps, np.exp(xb), rtol=1e-14)


def test_cpu_fused_matches_oracle():
    from tests.helpers import synth_diag_gauss

    D, N = 128, 64
    m, s, Minv, th, r = synth_diag_gauss(D, N, 20260923)
    model, metric = oc.Model(oc.DIAG_GAUSS, D, m, s), oc.Metric(oc.DIAG, Minv)
    z = oc.phasepoint(model, metric, th, r)
    a, _, _ = oc.leapfrog(model, metric, 0.1, z, 32)
    b = oc.leapfrog_omp(model, metric, 0.1, z, 32, n_threads=2)
    for f in ("theta", "r", "lp_gradient", "lp_value", "lk_value"):
        assert rel_err(getattr(b, f), getattr(a, f)) < 1e-12, f


# ------------------------------------------------------------------------------------------------ NUTS oracle
def _nuts_run(D, N, eps, seed, model=None, metric=None, max_depth=10):
    rng = np.random.default_rng(seed)
    model = model or oc.Model(oc.STD_NORMAL, D)
    metric = metric or oc.Metric(oc.UNIT)
    th = rng.normal(size=(D, N))
    z0 = oc.phasepoint(model, metric, th, np.zeros((D, N)))
    nt = rng.normal(size=(D, N))
    dirs = rng.integers(0, 2, size=(N, max_depth + 1)).astype(np.uint8)
    exps = rng.exponential(size=(N, 1 << max_depth))
    return oc.nuts_transition(model, metric, eps, z0, nt, dirs, exps, max_depth=max_depth), (th, nt, dirs, exps)


def test_nuts_oracle_determinism_and_structure():
    """test/trajectory.jl:125-141 (same randomness -> same transition) + tree bookkeeping invariants."""
    (z1, s1, u1), _ = _nuts_run(5, 40, 0.3, 1)
    (z2, s2, u2), _ = _nuts_run(5, 40, 0.3, 1)
    assert (z1.theta == z2.theta).all() and (s1.tree_depth == s2.tree_depth).all() and (u1 == u2).all()
    # a transition of depth j that ended by a U-turn of the whole tree has exactly 2^j - 1 leapfrog steps;
    # early termination inside the last subtree gives fewer.  Always: 2^(j) - 1 <= n_steps <= 2^(j+1) - 1
    j, n = s1.tree_depth, s1.n_steps
    assert ((n >= (1 << j) - 1) & (n <= (1 << (j + 1)) - 1)).all()
    # exponentials consumed: one per internal combine (post-order) + one per successful doubling
    assert (u1 <= n).all() and (u1 >= j).all()
    assert ((s1.acceptance_rate >= 0) & (s1.acceptance_rate <= 1)).all()
    assert (s1.is_accept == 1).all()
    assert np.allclose(s1.hamiltonian_energy, -(z1.lp_value + z1.lk_value))


def test_nuts_oracle_uturn_criterion_hand_formula():
    """test/trajectory.jl:249-325 analogue: for a depth-1 tree the generalised criterion is
    dot(rho, Minv r_left) <= 0 || dot(rho, Minv r_right) <= 0 with rho = r_left + r_right."""
    D = 4
    rng = np.random.default_rng(3)
    Minv = np.exp(rng.uniform(-1, 1, D))
    model, metric = oc.Model(oc.STD_NORMAL, D), oc.Metric(oc.DIAG, Minv)
    hits = 0
    for trial in range(200):
        th, r = rng.normal(size=(D, 1)), rng.normal(size=(D, 1))
        z0 = oc.phasepoint(model, metric, th, r)
        dirs = np.zeros((1, 3), dtype=np.uint8)  # always expand to the right
        (z1, st, used) = oc.nuts_transition(model, metric, 1.2, z0, None, dirs, np.full((1, 8), 1e9), max_depth=2)
        zr, _, _ = oc.leapfrog(model, metric, 1.2, z0, 1)
        rho = z0.r[:, 0] + zr.r[:, 0]
        turn = (rho @ (Minv * z0.r[:, 0]) <= 0) or (rho @ (Minv * zr.r[:, 0]) <= 0)
        # a U-turn after the first doubling stops the loop with tree_depth 1 and n_steps 1
        if turn:
            hits += 1
            assert st.tree_depth[0] == 1 and st.n_steps[0] == 1
        else:
            assert st.n_steps[0] > 1
    assert 10 < hits < 190


def test_nuts_oracle_samples_standard_normal():
    """statistical pin of the whole NUTS restatement (test/sampler.jl style): 1 chain x 3000 transitions."""
    D = 3
    rng = np.random.default_rng(4)
    model, metric = oc.Model(oc.STD_NORMAL, D), oc.Metric(oc.UNIT)
    th = np.zeros((D, 1))
    draws = []
    z = oc.phasepoint(model, metric, th, np.zeros((D, 1)))
    for it in range(3000):
        nt = rng.normal(size=(D, 1))
        dirs = rng.integers(0, 2, size=(1, 11)).astype(np.uint8)
        exps = rng.exponential(size=(1, 1024))
        z, st, _ = oc.nuts_transition(model, metric, 0.9, z, nt, dirs, exps)
        draws.append(z.theta[:, 0].copy())
    X = np.array(draws[200:])
    assert np.abs(X.mean(axis=0)).max() < 0.12 and np.abs(X.var(axis=0) - 1).max() < 0.15


def test_nuts_oracle_classic_and_strict_hand_formulas():
    """First doubling, expanding right from z0 to z1 (tleft = leaf z0, tright = leaf z1):
    ClassicNoUTurn (trajectory.jl:551-557): dot(dth, Minv(-r0)) >= 0 || dot(-dth, Minv r1) >= 0, dth = th1 - th0.
    StrictGeneralisedNoUTurn (:579-613) on two single leaves reduces to the generalised check on rho = r0 + r1
    (both sub-tree checks use the same rho and the same two momenta)."""
    D = 4
    rng = np.random.default_rng(13)
    Minv = np.exp(rng.uniform(-1, 1, D))
    model, metric = oc.Model(oc.STD_NORMAL, D), oc.Metric(oc.DIAG, Minv)
    hits = 0
    for trial in range(200):
        th, r = rng.normal(size=(D, 1)), rng.normal(size=(D, 1))
        z0 = oc.phasepoint(model, metric, th, r)
        dirs = np.zeros((1, 3), dtype=np.uint8)
        big = np.full((1, 8), 1e9)
        zr, _, _ = oc.leapfrog(model, metric, 1.2, z0, 1)
        dth = zr.theta[:, 0] - z0.theta[:, 0]
        turn_c = (dth @ (Minv * -z0.r[:, 0]) >= 0) or (-dth @ (Minv * zr.r[:, 0]) >= 0)
        _, st, _ = oc.nuts_transition(model, metric, 1.2, z0, None, dirs, big, max_depth=2, criterion="classic")
        assert (st.n_steps[0] == 1) == bool(turn_c)
        hits += bool(turn_c)
        _, sg, _ = oc.nuts_transition(model, metric, 1.2, z0, None, dirs, big, max_depth=1)
        _, ss, _ = oc.nuts_transition(model, metric, 1.2, z0, None, dirs, big, max_depth=1, criterion="strict")
        assert sg.n_steps[0] == ss.n_steps[0] == 1
    assert 5 < hits < 195


@pytest.mark.parametrize("sampler,criterion", [("slice", "generalised"), ("multinomial", "classic"),
                                               ("multinomial", "strict"), ("slice", "strict")])
def test_nuts_oracle_variants_sample_standard_normal(sampler, criterion):
    """every (trajectory sampler, criterion) pair leaves N(0, I) invariant; strict never grows a larger tree than
    generalised from the same randomness."""
    D = 3
    rng = np.random.default_rng(5)
    model, metric = oc.Model(oc.STD_NORMAL, D), oc.Metric(oc.UNIT)
    z = oc.phasepoint(model, metric, np.zeros((D, 1)), np.zeros((D, 1)))
    draws = []
    for it in range(2500):
        nt = rng.normal(size=(D, 1))
        dirs = rng.integers(0, 2, size=(1, 11)).astype(np.uint8)
        exps = rng.exponential(size=(1, 1024))
        if sampler == "slice":
            exps[:, 1:] = rng.uniform(size=(1, 1023))
        if criterion == "strict" and sampler == "multinomial":
            _, sg, _ = oc.nuts_transition(model, metric, 0.9, z, nt, dirs, exps)
        z, st, _ = oc.nuts_transition(model, metric, 0.9, z, nt, dirs, exps, sampler=sampler, criterion=criterion)
        if criterion == "strict" and sampler == "multinomial":
            assert st.n_steps[0] <= sg.n_steps[0]
        draws.append(z.theta[:, 0].copy())
    X = np.array(draws[200:])
    assert np.abs(X.mean(axis=0)).max() < 0.15 and np.abs(X.var(axis=0) - 1).max() < 0.18


# ------------------------------------------------------------------------------------------------ NUTS known answers
from tests.helpers import nuts_golden_cases  # noqa: E402

_NUTS_GOLD = nuts_golden_cases()


@pytest.mark.parametrize("case", _NUTS_GOLD, ids=[c["name"] for c in _NUTS_GOLD])
def test_nuts_oracle_matches_mp50_recursive_restatement(case):
    """The C oracle's NUTS transition against tests/golden/nuts_mp50.json -- an independent recursive restatement of
    src/trajectory.jl:626-742 evaluated in 50-digit arithmetic (tests/golden/gen_nuts_mp.py): same tree for every
    chain (depth, leapfrog steps, divergence flag, number of random variates consumed), same selected candidate and
    statistics to 1e-10, for both trajectory samplers and all three termination criteria."""
    D, N = case["D"], case["N"]
    kinds = dict(std_normal=oc.STD_NORMAL, diag_gauss=oc.DIAG_GAUSS, dense_gauss=oc.DENSE_GAUSS, funnel=oc.FUNNEL)
    mkinds = dict(unit=oc.UNIT, diag=oc.DIAG, dense=oc.DENSE)
    p0 = None if case["p0"] is None else np.array(case["p0"])
    p1 = None if case["p1"] is None else np.asfortranarray(np.array(case["p1"]))
    Minv = None if case["Minv"] is None else np.asfortranarray(np.array(case["Minv"]))
    model, metric = oc.Model(kinds[case["model"]], D, p0, p1, case["c0"]), oc.Metric(mkinds[case["metric"]], Minv)
    th, r = np.array(case["theta0"]).T, np.array(case["r0"]).T
    z0 = oc.phasepoint(model, metric, th, r)
    oc.set_tempering(case.get("temper_alpha", 0.0))  # > 0: every leaf is a TemperedLeapfrog step
    try:
        z, st, used = oc.nuts_transition(model, metric, case["eps"], z0, None, np.array(case["dirs"], dtype=np.uint8),
                                         np.array(case["variates"]), max_depth=case["max_depth"], delta_max=case["delta_max"],
                                         sampler=case["sampler"], criterion=case["criterion"])
    finally:
        oc.set_tempering(0.0)
    e = case["expect"]
    assert (st.tree_depth == np.array(e["tree_depth"])).all()
    assert (st.n_steps == np.array(e["n_steps"])).all()
    assert (st.numerical_error.astype(bool) == np.array(e["numerical_error"])).all()
    assert (used == np.array(e["variates_used"])).all()
    assert rel_err(z.theta, np.array(e["theta"]).T) < 1e-10 and rel_err(z.r, np.array(e["r"]).T) < 1e-10
    assert rel_err(z.lp_gradient, np.array(e["lp_gradient"]).T) < 1e-10
    assert np.allclose(z.lp_value, e["lp_value"], rtol=1e-10, atol=1e-10) and np.allclose(z.lk_value, e["lk_value"], rtol=1e-10, atol=1e-10)
    assert np.allclose(st.acceptance_rate, e["acceptance_rate"], rtol=1e-10)
    assert np.allclose(st.hamiltonian_energy_error, e["hamiltonian_energy_error"], rtol=1e-9, atol=1e-10)
    assert np.allclose(st.max_hamiltonian_energy_error, e["max_hamiltonian_energy_error"], rtol=1e-9, atol=1e-10)


from tests.helpers import hmc_golden_cases  # noqa: E402

_HMC_GOLD = hmc_golden_cases()


@pytest.mark.parametrize("case", _HMC_GOLD, ids=[c["name"] for c in _HMC_GOLD])
def test_hmc_oracle_matches_mp50_restatement(case):
    """The C oracle's static transitions (momentum refresh from normals, EndPointTS Metropolis step with momentum flip,
    MultinomialTS trajectory sampling) against tests/golden/hmc_mp50.json (50-digit restatement, gen_hmc_mp.py)."""
    D, N = case["D"], case["N"]
    kinds = dict(std_normal=oc.STD_NORMAL, diag_gauss=oc.DIAG_GAUSS, dense_gauss=oc.DENSE_GAUSS, funnel=oc.FUNNEL)
    mkinds = dict(unit=oc.UNIT, diag=oc.DIAG, dense=oc.DENSE)
    p0 = None if case["p0"] is None else np.array(case["p0"])
    p1 = None if case["p1"] is None else np.asfortranarray(np.array(case["p1"]))
    Minv = None if case["Minv"] is None else np.asfortranarray(np.array(case["Minv"]))
    model, metric = oc.Model(kinds[case["model"]], D, p0, p1, case["c0"]), oc.Metric(mkinds[case["metric"]], Minv)
    th, nt = np.array(case["theta0"]).T, np.array(case["normals"]).T
    z0 = oc.phasepoint(model, metric, th, np.zeros((D, N)))
    oc.set_tempering(case.get("temper_alpha", 0.0))  # > 0: TemperedLeapfrog(eps, alpha) is the transition's integrator
    try:
        if case["sampler"] == "endpoint":
            z, st = oc.hmc_transition(model, metric, case["eps"], case["n_steps"], z0, nt, np.array(case["variates"]))
        else:
            z, st = oc.hmc_multinomial_transition(model, metric, case["eps"], case["n_steps"], case["n_fwd"], z0, nt,
                                                  np.array(case["variates"]))
            if "index" in case["expect"]:
                assert (st.tree_depth == np.array(case["expect"]["index"])).all()
    finally:
        oc.set_tempering(0.0)
    e = case["expect"]
    assert (st.is_accept.astype(bool) == np.array(e["is_accept"])).all()
    assert rel_err(z.theta, np.array(e["theta"]).T) < 1e-10 and rel_err(z.r, np.array(e["r"]).T) < 1e-10
    assert rel_err(z.lp_gradient, np.array(e["lp_gradient"]).T) < 1e-10
    assert np.allclose(z.lp_value, e["lp_value"], rtol=1e-10, atol=1e-10) and np.allclose(z.lk_value, e["lk_value"], rtol=1e-10, atol=1e-10)
    assert np.allclose(st.acceptance_rate, e["acceptance_rate"], rtol=1e-10)
    assert np.allclose(st.hamiltonian_energy_error, e["hamiltonian_energy_error"], rtol=1e-9, atol=1e-10)


# ------------------------------------------------------------------------------------------------ structural properties
def _systems():
    rng = np.random.default_rng(21)
    D = 4
    B = rng.normal(size=(D, D))
    P = B @ B.T / D + np.eye(D)
    C_ = rng.normal(size=(D, D))
    Minv_dense = C_ @ C_.T / D + 0.5 * np.eye(D)
    return [
        ("diag_gauss/diag", oc.Model(oc.DIAG_GAUSS, D, rng.normal(size=D), np.exp(rng.uniform(-0.5, 0.5, D))), oc.Metric(oc.DIAG, np.exp(rng.uniform(-0.5, 0.5, D)))),
        ("dense_gauss/dense", oc.Model(oc.DENSE_GAUSS, D, rng.normal(size=D), np.asfortranarray(P)), oc.Metric(oc.DENSE, np.asfortranarray(Minv_dense))),
        ("funnel/unit", oc.Model(oc.FUNNEL, D), oc.Metric(oc.UNIT)),
        ("funnel/dense", oc.Model(oc.FUNNEL, D), oc.Metric(oc.DENSE, np.asfortranarray(Minv_dense))),
    ]


@pytest.mark.parametrize("name,model,metric", _systems(), ids=[s[0] for s in _systems()])
def test_leapfrog_map_is_symplectic_and_volume_preserving(name, model, metric):
    """The leapfrog map is a composition of shears, hence symplectic for ANY target and metric: its Jacobian J satisfies
    J' Omega J = Omega (and det J = 1).  Checked by central differences on the restated `step` -- an error in the
    order / sign / scaling of a kick or drift breaks it at O(1), independently of any known answer."""
    D = 4
    rng = np.random.default_rng(5)
    z0 = np.concatenate([rng.normal(size=D) * 0.5, rng.normal(size=D)])
    eps, n = 0.13, 3

    def f(v):
        z = oc.phasepoint(model, metric, v[:D, None].copy(), v[D:, None].copy())
        z1 = oc.leapfrog(model, metric, eps, z, n)[0]
        return np.concatenate([z1.theta[:, 0], z1.r[:, 0]])

    hstep = 1e-5
    J = np.zeros((2 * D, 2 * D))
    for k in range(2 * D):
        e = np.zeros(2 * D)
        e[k] = hstep
        J[:, k] = (f(z0 + e) - f(z0 - e)) / (2 * hstep)
    Om = np.block([[np.zeros((D, D)), np.eye(D)], [-np.eye(D), np.zeros((D, D))]])
    assert np.abs(J.T @ Om @ J - Om).max() < 1e-7
    assert abs(np.linalg.det(J) - 1.0) < 1e-7


@pytest.mark.parametrize("name,model,metric", _systems(), ids=[s[0] for s in _systems()])
def test_energy_error_is_second_order_in_the_step_size(name, model, metric):
    """Over a fixed integration time the energy error of leapfrog scales like eps^2: halving eps divides it by ~4."""
    D = 4
    rng = np.random.default_rng(6)
    th, r = rng.normal(size=(D, 6)) * 0.4, rng.normal(size=(D, 6))
    z = oc.phasepoint(model, metric, th, r)
    H0 = z.energy()
    errs = []
    for eps, n in ((0.04, 8), (0.02, 16), (0.01, 32)):
        z1 = oc.leapfrog(model, metric, eps, z, n)[0]
        errs.append(np.abs(z1.energy() - H0))
    r1, r2 = errs[0] / errs[1], errs[1] / errs[2]
    assert np.all((r1 > 2.5) & (r1 < 6.5)) and np.all((r2 > 3.0) & (r2 < 5.5)), (r1, r2)


def test_gaussian_leapfrog_map_is_linear():
    """Gaussian target (mean 0) + Euclidean metric: the n-step map is linear in (theta, r) -- the property the fused
    fast path and the tiled DMMA kernel rely on (magnitude proof, energies evaluated once at the end)."""
    rng = np.random.default_rng(7)
    D = 5
    B = rng.normal(size=(D, D))
    model = oc.Model(oc.DENSE_GAUSS, D, np.zeros(D), np.asfortranarray(B @ B.T / D + np.eye(D)))
    metric = oc.Metric(oc.DIAG, np.exp(rng.uniform(-0.5, 0.5, D)))

    def f(th, r):
        z1 = oc.leapfrog(model, metric, 0.11, oc.phasepoint(model, metric, th[:, None].copy(), r[:, None].copy()), 9)[0]
        return z1.theta[:, 0], z1.r[:, 0]

    a, b = 0.7, -1.9
    t1, r1, t2, r2 = (rng.normal(size=D) for _ in range(4))
    A1, A2, A3 = f(t1, r1), f(t2, r2), f(a * t1 + b * t2, a * r1 + b * r2)
    assert np.allclose(A3[0], a * A1[0] + b * A2[0], rtol=1e-12, atol=1e-12)
    assert np.allclose(A3[1], a * A1[1] + b * A2[1], rtol=1e-12, atol=1e-12)


# ------------------------------------------------------------------------------------------------ the kernel's iterative scheme
def _np_system(kind, D, p0, p1, mkind, Minv, eps):
    """numpy closures for oracle/nuts_iterative.py from fixture-style parameters."""
    from oracle import nuts_iterative as ni

    p0 = None if p0 is None else np.asarray(p0, dtype=np.float64)
    p1 = None if p1 is None else np.asarray(p1, dtype=np.float64)
    Minv = None if Minv is None else np.asarray(Minv, dtype=np.float64)

    def logp_grad(th):
        if kind == "std_normal":
            return -0.5 * float(th @ th), th.copy()
        if kind == "diag_gauss":
            d = th - p0
            return -0.5 * float(np.sum(d * d / (p1 * p1))), d / (p1 * p1)
        if kind == "dense_gauss":
            d = th - p0
            Pd = p1 @ d
            return -0.5 * float(d @ Pd), Pd
        v, x = th[0], th[1:]
        ev = math.exp(-v)
        S = float(np.sum(x * x) * ev)
        return -v * v / 18 - 0.5 * (S + (D - 1) * v), np.concatenate([[v / 9 - 0.5 * (S - (D - 1))], x * ev])

    def dHdr(r):
        if mkind == "unit":
            return r
        return Minv * r if mkind == "diag" else Minv @ r

    return ni.System(logp_grad, dHdr, eps)


@pytest.mark.parametrize("case", _NUTS_GOLD, ids=[c["name"] for c in _NUTS_GOLD])
def test_iterative_nuts_scheme_matches_mp50_recursion(case):
    """oracle/nuts_iterative.py -- the binary-counter / pending-level / float-up scheme the CUDA kernel runs, written out
    in numpy -- reproduces the recursive 50-digit fixtures: the iterative algorithm is pinned on the CPU."""
    from oracle import nuts_iterative as ni

    S = _np_system(case["model"], case["D"], case["p0"], case["p1"], case["metric"], case["Minv"], case["eps"])
    S.temper_alpha = case.get("temper_alpha", 0.0)
    e = case["expect"]
    for c in range(case["N"]):
        z0 = S.point(np.array(case["theta0"][c]), np.array(case["r0"][c]))
        zc, st, used = ni.transition(S, z0, case["dirs"][c], case["variates"][c], sampler=case["sampler"],
                                     criterion=case["criterion"], max_depth=case["max_depth"], delta_max=case["delta_max"])
        assert (st["tree_depth"], st["n_steps"], bool(st["numerical_error"]), used) == \
               (e["tree_depth"][c], e["n_steps"][c], e["numerical_error"][c], e["variates_used"][c]), c
        assert rel_err(zc["th"], e["theta"][c]) < 1e-10 and rel_err(zc["r"], e["r"][c]) < 1e-10
        assert st["acceptance_rate"] == pytest.approx(e["acceptance_rate"][c], rel=1e-10)
        assert st["max_hamiltonian_energy_error"] == pytest.approx(e["max_hamiltonian_energy_error"][c], rel=1e-9, abs=1e-10)
        if case["sampler"] == "multinomial":  # the log-free (m, w) weight form staged for the kernel decides the same
            zm, sm, um = ni.transition(S, z0, case["dirs"][c], case["variates"][c], sampler="multinomial",
                                       criterion=case["criterion"], max_depth=case["max_depth"], delta_max=case["delta_max"],
                                       max_weights=True)
            assert (sm["tree_depth"], sm["n_steps"], um) == (e["tree_depth"][c], e["n_steps"][c], e["variates_used"][c])
            assert np.array_equal(zm["th"], zc["th"])


@pytest.mark.parametrize("sampler,criterion", [("multinomial", "generalised"), ("slice", "generalised"),
                                               ("multinomial", "classic"), ("multinomial", "strict"), ("slice", "strict")])
@pytest.mark.parametrize("eps,delta_max", [(0.12, 6.0), (0.4, 0.3)], ids=["deep", "divergent"])
def test_iterative_nuts_scheme_matches_recursive_c_oracle_on_deep_trees(sampler, criterion, eps, delta_max):
    """Same check against the (recursive) C oracle on 120 random chains with trees up to depth 8, small step sizes and
    a tight Delta_max so that max-depth exits, U-turns inside subtrees (float-up) and divergences all occur."""
    from oracle import nuts_iterative as ni

    rng = np.random.default_rng(31)
    D, N, max_depth = 5, 120, 8
    sd = np.exp(rng.uniform(-1.0, 1.0, D))
    mu = rng.normal(size=D)
    Minv = np.exp(rng.uniform(-0.5, 0.5, D))
    model, metric = oc.Model(oc.DIAG_GAUSS, D, mu, sd, 0.0), oc.Metric(oc.DIAG, Minv)
    S = _np_system("diag_gauss", D, mu, sd, "diag", Minv, eps)
    th, r = rng.normal(size=(D, N)) * 2.0, rng.normal(size=(D, N))
    dirs = rng.integers(0, 2, size=(N, max_depth + 1)).astype(np.uint8)
    var = rng.exponential(size=(N, 1 << max_depth))
    if sampler == "slice":
        var[:, 1:] = rng.uniform(size=(N, (1 << max_depth) - 1))
    z0 = oc.phasepoint(model, metric, th, r)
    zo, so, used = oc.nuts_transition(model, metric, eps, z0, None, dirs, var, max_depth=max_depth, delta_max=delta_max,
                                      sampler=sampler, criterion=criterion)
    depths = set()
    for c in range(N):
        zc, st, nu = ni.transition(S, S.point(th[:, c].copy(), r[:, c].copy()), dirs[c], var[c], sampler=sampler,
                                   criterion=criterion, max_depth=max_depth, delta_max=delta_max)
        assert (st["tree_depth"], st["n_steps"], int(st["numerical_error"]), nu) == \
               (so.tree_depth[c], so.n_steps[c], so.numerical_error[c], used[c]), c
        assert rel_err(zc["th"], zo.theta[:, c]) < 1e-9
        depths.add(st["tree_depth"])
        if sampler == "multinomial":  # the probability-domain form of the same decisions (planned kernel optimisation)
            zl, sl, nl = ni.transition(S, S.point(th[:, c].copy(), r[:, c].copy()), dirs[c], var[c], sampler=sampler,
                                       criterion=criterion, max_depth=max_depth, delta_max=delta_max, linear_accept=True)
            assert (sl["tree_depth"], sl["n_steps"], nl) == (st["tree_depth"], st["n_steps"], nu)
            assert np.array_equal(zl["th"], zc["th"])
            zm, sm, nm = ni.transition(S, S.point(th[:, c].copy(), r[:, c].copy()), dirs[c], var[c], sampler=sampler,
                                       criterion=criterion, max_depth=max_depth, delta_max=delta_max, max_weights=True)
            assert (sm["tree_depth"], sm["n_steps"], nm) == (st["tree_depth"], st["n_steps"], nu)
            assert np.array_equal(zm["th"], zc["th"])  # (m, w) weights: same decisions, no log
    if delta_max < 1.0:
        assert so.numerical_error.sum() > 5       # numerical terminations inside and at the top of subtrees
    else:
        assert len(depths) >= 3                   # several tree sizes, partial subtrees (float-up) included
        assert (((so.n_steps + 1) & so.n_steps) != 0).any()


def test_iterative_scheme_vs_recursive_oracle_randomised_configurations():
    """Differential test over random configurations (dimension, target, metric incl. dense, step size, depth limit,
    divergence threshold, sampler, criterion): the iterative numpy scheme and the recursive C oracle must build the same
    tree and select the same point for every chain."""
    from oracle import nuts_iterative as ni

    rng = np.random.default_rng(77)
    n_cfg, checked = 60, 0
    for cfg in range(n_cfg):
        D = int(rng.integers(1, 9))
        kind = ["std_normal", "diag_gauss", "dense_gauss", "funnel"][int(rng.integers(0, 4))]
        if kind == "funnel" and D < 2:
            D = 2
        mkind = ["unit", "diag", "dense"][int(rng.integers(0, 3))]
        eps = float(np.exp(rng.uniform(np.log(0.05), np.log(0.8))))
        max_depth = int(rng.integers(1, 8))
        delta_max = [1000.0, 2.0][int(rng.integers(0, 2))]
        sampler = ["multinomial", "slice"][int(rng.integers(0, 2))]
        criterion = ["generalised", "classic", "strict"][int(rng.integers(0, 3))]
        p0 = p1 = Minv = None
        if kind == "diag_gauss":
            p0, p1 = rng.normal(size=D), np.exp(rng.uniform(-0.7, 0.7, D))
        elif kind == "dense_gauss":
            B = rng.normal(size=(D, D))
            p0, p1 = rng.normal(size=D), B @ B.T / D + np.eye(D)
        if mkind == "diag":
            Minv = np.exp(rng.uniform(-0.7, 0.7, D))
        elif mkind == "dense":
            B = rng.normal(size=(D, D))
            Minv = B @ B.T / D + 0.5 * np.eye(D)
        kinds = dict(std_normal=oc.STD_NORMAL, diag_gauss=oc.DIAG_GAUSS, dense_gauss=oc.DENSE_GAUSS, funnel=oc.FUNNEL)
        mkinds = dict(unit=oc.UNIT, diag=oc.DIAG, dense=oc.DENSE)
        model = oc.Model(kinds[kind], D, p0, None if p1 is None else np.asfortranarray(p1), 0.0)
        metric = oc.Metric(mkinds[mkind], None if Minv is None else np.asfortranarray(Minv))
        S = _np_system(kind, D, p0, p1, mkind, Minv, eps)
        N = 12
        th, r = rng.normal(size=(D, N)) * (0.6 if kind == "funnel" else 1.5), rng.normal(size=(D, N))
        dirs = rng.integers(0, 2, size=(N, max_depth + 1)).astype(np.uint8)
        var = rng.exponential(size=(N, 1 << max_depth))
        if sampler == "slice":
            var[:, 1:] = rng.uniform(size=(N, (1 << max_depth) - 1))
        zo, so, used = oc.nuts_transition(model, metric, eps, oc.phasepoint(model, metric, th, r), None, dirs, var,
                                          max_depth=max_depth, delta_max=delta_max, sampler=sampler, criterion=criterion)
        for c in range(N):
            zc, st, nu = ni.transition(S, S.point(th[:, c].copy(), r[:, c].copy()), dirs[c], var[c], sampler=sampler,
                                       criterion=criterion, max_depth=max_depth, delta_max=delta_max)
            key = (cfg, c, kind, mkind, sampler, criterion)
            if not np.all(np.isfinite(zo.theta[:, c])):
                continue  # a non-finite trajectory: covered by the -Inf mapping tests, not a tree-logic question
            assert (st["tree_depth"], st["n_steps"], int(st["numerical_error"]), nu) == \
                   (so.tree_depth[c], so.n_steps[c], so.numerical_error[c], used[c]), key
            assert rel_err(zc["th"], zo.theta[:, c]) < 1e-8, key
            checked += 1
    assert checked > 0.9 * n_cfg * 12
