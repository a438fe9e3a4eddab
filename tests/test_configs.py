"""GPU end-to-end checks of the five BASELINE.json configurations at their own dimensions (C2 / C3 at their own 4096 chains,
C4 / C5 at one GPU's share or less of the chains): the sampling loop
(src/sampler.jl:159-248 mirror) with pooled adaptation on top of the fused transition kernels recovers the
targets' moments.  (Trajectory-level parity with the oracle is in test_gpu_parity.py.)"""
import numpy as np
import pytest
import torch

import ahmc_b200 as A
from ahmc_b200 import adaptation as ad
from tests.helpers import rel_err

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _corr_gauss(D, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    Q, _ = np.linalg.qr(rng.normal(size=(D, D)))
    lam = np.exp(np.linspace(np.log(0.1), np.log(10.0), D))
    return (Q * lam) @ Q.T, (Q / lam) @ Q.T


def test_c1_static_hmc_unit_metric_d10_64_chains():
    """StaticTrajectory(Leapfrog(0.1), 32) + UnitEuclideanMetric, D=10 std-Normal, 64 chains (test/sampler-vec.jl path).
    Integration time 3.2 ~ pi makes this config nearly anti-periodic (theta' ~ -0.998 theta - 0.058 r): the lag-1
    autocorrelation must be ~ cos(3.2) and the stationary variance 1 is reached only over ~250 transitions, so the
    check runs 6000 transitions in one persistent launch."""
    D, N, T = 10, 64, 6000
    h = A.Hamiltonian(A.UnitEuclideanMetric(D), A.StdNormal(D))
    kern = A.HMCKernel(A.Trajectory(A.EndPointTS, A.Leapfrog(0.1), A.FixedNSteps(32)))
    th0 = torch.rand((N, D), dtype=torch.float64, device=DEV)
    zl, draws, st = A.sample_transitions(A.PhiloxRNG(1), h, kern, A.phasepoint(h, th0, torch.zeros_like(th0)), T)
    X = draws[1000:].cpu().numpy()  # (T', N, D)
    assert st["acceptance_rate"].mean().item() > 0.99
    lag1 = np.mean(X[1:] * X[:-1]) / np.mean(X * X)
    assert abs(lag1 - np.cos(3.2)) < 0.01
    assert np.abs(X.reshape(-1, D).mean(axis=0)).max() < 0.05
    assert abs(X.var() - 1) < 0.25  # RNDATOL-style tolerance (test/common.jl:12)


def test_c2_hmcda_diag_metric_correlated_gaussian():
    """HMCDA(0.8, lambda=1) + DiagEuclideanMetric on a D=128 correlated Gaussian; shared eps from pooled dual averaging."""
    D, N = 128, 4096
    Sigma, P = _corr_gauss(D, 7)
    h = A.Hamiltonian(A.DiagEuclideanMetric(np.diag(Sigma).copy()), A.DenseGaussian(np.zeros(D), P))
    kern = A.HMCKernel(A.Trajectory(A.EndPointTS, A.Leapfrog(0.05), A.FixedIntegrationTime(1.0)))
    adaptor = ad.NaiveHMCAdaptor(ad.UnitMassMatrix(), ad.NesterovDualAveraging(0.8, 0.05))
    th0 = torch.as_tensor(np.random.default_rng(0).normal(size=(N, D)), device=DEV)
    res = ad.sample(A.PhiloxRNG(2), h, kern, th0, 150, adaptor, 100, keep_draws=True, drop_warmup=True)
    acc = np.mean([s["acceptance_rate"] for s in res.stats[100:]])
    assert 0.6 < acc < 0.95, acc          # dual averaging steers the pooled acceptance towards delta = 0.8
    assert 0.02 < res.eps < 1.0
    X = torch.stack(res.draws).reshape(-1, D).cpu().numpy()
    assert np.abs(X.mean(axis=0)).max() < 0.2
    ratio = X.var(axis=0) / np.diag(Sigma)
    assert 0.7 < ratio.min() and ratio.max() < 1.35


def test_c3_nuts_diag_metric_d128():
    D, N = 128, 4096
    s = np.exp(np.linspace(np.log(0.1), np.log(10.0), D))
    h = A.Hamiltonian(A.DiagEuclideanMetric(s * s), A.DiagGaussian(np.zeros(D), s))
    kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.25), A.GeneralisedNoUTurn()))
    z = A.phasepoint(h, torch.zeros((N, D), dtype=torch.float64, device=DEV), torch.zeros((N, D), dtype=torch.float64, device=DEV))
    zl, draws, st = A.sample_transitions(A.PhiloxRNG(3), h, kern, z, 30)
    X = draws[10:].reshape(-1, D).cpu().numpy() / s
    assert np.abs(X.mean(axis=0)).max() < 0.08 and np.abs(X.var(axis=0) - 1).max() < 0.12
    assert st["acceptance_rate"].mean().item() > 0.8 and int(st["numerical_error"].sum().item()) == 0


def test_c4_funnel_nuts_stan_adaptor_pooled():
    """C4 at BASELINE's own dimension: NUTS + StanHMCAdaptor on Neal's funnel D=100, pooled windows, the adaptor resident on
    the device (ahmc_adapt_exchange_f64: no host synchronisation during warm-up): v ~ N(0, 3^2)."""
    D, N = 100, 1024
    h = A.Hamiltonian(A.DiagEuclideanMetric(np.ones(D)), A.Funnel(D))
    kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.1), A.GeneralisedNoUTurn()))
    th0 = torch.as_tensor(np.random.default_rng(1).normal(size=(N, D)) * 0.1, device=DEV)
    res = ad.sample_pooled_device(A.PhiloxRNG(4), h, kern, th0, 200, 200, eps0=0.1)
    assert res.Minv is not None and res.Minv.shape == (D,) and res.Minv[0] > 1.0  # adapted: var(v) >> var at init
    assert 0.01 < res.eps < 2.0
    # draws with the adapted eps / M^-1 (persistent launch, kept)
    hd = A.Hamiltonian(A.DiagEuclideanMetric(res.Minv), A.Funnel(D))
    kd = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(res.eps), A.GeneralisedNoUTurn()))
    z = A.phasepoint(hd, res.theta, torch.zeros_like(res.theta))
    zl, draws, st = A.sample_transitions(A.PhiloxRNG(9), hd, kd, z, 60)
    v = draws[:, :, 0].reshape(-1).cpu().numpy()
    # the funnel's neck is famously under-sampled by NUTS; the bulk of v ~ N(0, 9) must be there
    assert abs(np.median(v)) < 1.0 and 1.5 < v.std() < 3.6


def test_c4_funnel_host_pooled_adaptor_small():
    """the host-side pooled loop (`sample`, adaptation.py) on a small funnel -- the path the gloo tests exercise"""
    D, N = 20, 512
    h = A.Hamiltonian(A.DiagEuclideanMetric(np.ones(D)), A.Funnel(D))
    kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.1), A.GeneralisedNoUTurn()))
    adaptor = ad.StanHMCAdaptor(ad.WelfordVar(D), ad.NesterovDualAveraging(0.8, 0.1), init_buffer=20, term_buffer=20, window_size=10)
    th0 = torch.as_tensor(np.random.default_rng(1).normal(size=(N, D)) * 0.1, device=DEV)
    res = ad.sample(A.PhiloxRNG(4), h, kern, th0, 160, adaptor, 120, keep_draws=True, drop_warmup=True)
    v = torch.stack(res.draws)[:, :, 0].reshape(-1).cpu().numpy()
    assert abs(np.median(v)) < 1.2 and 1.2 < v.std() < 3.8
    assert res.Minv is not None and res.Minv[0] > 1.0 and 0.01 < res.eps < 2.0


def test_c5_nuts_dense_metric_d256():
    """NUTS + DenseEuclideanMetric (Minv = Sigma) on a D=256 correlated Gaussian: with the exact metric the
    sampler sees an isotropic problem (1024 chains = one GPU's share of C5's 8192)."""
    D, N = 256, 1024
    Sigma, P = _corr_gauss(D, 11)
    h = A.Hamiltonian(A.DenseEuclideanMetric(Sigma), A.DenseGaussian(np.zeros(D), P))
    kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.4), A.GeneralisedNoUTurn()))
    z = A.phasepoint(h, torch.zeros((N, D), dtype=torch.float64, device=DEV), torch.zeros((N, D), dtype=torch.float64, device=DEV))
    zl, draws, st = A.sample_transitions(A.PhiloxRNG(5), h, kern, z, 25)
    X = draws[5:].reshape(-1, D).cpu().numpy()
    L = np.linalg.cholesky(Sigma)
    W = np.linalg.solve(L, X.T).T  # whitened draws ~ N(0, I)
    assert np.abs(W.mean(axis=0)).max() < 0.2 and abs(W.var(axis=0).mean() - 1) < 0.05
    assert st["acceptance_rate"][5:].mean().item() > 0.6
    # (with the exact metric every mode has the same frequency: NUTS trees resonate and grow deep -- not asserted)


def test_nuts_dense_metric_adapted_with_pooled_welford_cov():
    """NUTS + DenseEuclideanMetric whose M^-1 is adapted by the pooled WelfordCov (K5 + K5b records): the adapted
    metric approaches the target covariance and the draws recover it (8a a15, massmatrix.jl:286-340)."""
    D, N = 16, 512
    Sigma, P = _corr_gauss(D, 11)
    h = A.Hamiltonian(A.DenseEuclideanMetric(np.eye(D)), A.DenseGaussian(np.zeros(D), P))
    kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.1), A.GeneralisedNoUTurn()))
    adaptor = ad.StanHMCAdaptor(ad.WelfordCov(D), ad.NesterovDualAveraging(0.8, 0.1))
    th0 = torch.as_tensor(np.random.default_rng(1).normal(size=(N, D)), device=DEV)
    res = ad.sample(A.PhiloxRNG(5), h, kern, th0, 200, adaptor, 150, keep_draws=True, drop_warmup=True)
    assert res.Minv.shape == (D, D)
    rel = np.linalg.norm(res.Minv - Sigma) / np.linalg.norm(Sigma)
    assert rel < 0.15, rel
    X = torch.stack(list(res.draws)).reshape(-1, D).cpu().numpy()
    emp = np.cov(X.T)
    assert np.linalg.norm(emp - Sigma) / np.linalg.norm(Sigma) < 0.15
    acc = np.mean([s["acceptance_rate"] for s in res.stats[150:]])
    assert 0.6 < acc < 0.97, acc
    assert res.timing["sampling_launch"] > 0  # the 50 post-warm-up transitions ran as one persistent launch


def test_nuts_diag_metric_adapted_with_pooled_nutpie_var():
    """NutpieVar (positions + gradients, massmatrix.jl:172-250): for an independent Gaussian N(0, s^2) the gradient
    is -x/s^2, so sqrt(var_x / var_g) = s^2 exactly: the adapted M^-1 recovers the target variances."""
    D, N = 24, 512
    sd = np.exp(np.linspace(np.log(0.2), np.log(5.0), D))
    h = A.Hamiltonian(A.DiagEuclideanMetric(np.ones(D)), A.DiagGaussian(np.zeros(D), sd))
    kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.1), A.GeneralisedNoUTurn()))
    adaptor = ad.StanHMCAdaptor(ad.NutpieVar(D), ad.NesterovDualAveraging(0.8, 0.1))
    th0 = torch.as_tensor(np.random.default_rng(2).normal(size=(N, D)), device=DEV)
    res = ad.sample(A.PhiloxRNG(6), h, kern, th0, 180, adaptor, 150)  # 150 warm-ups: one metric window ending at 100
    ratio = res.Minv / sd ** 2
    assert 0.8 < ratio.min() and ratio.max() < 1.25, (ratio.min(), ratio.max())
    acc = np.mean([s["acceptance_rate"] for s in res.stats[150:]])
    assert 0.6 < acc < 0.97, acc


def test_in_launch_per_chain_adaptation_equals_host_loop_with_oracle_adaptors():
    """ahmc_nuts_adapt_sample_f64 (per-chain NesterovDualAveraging + windowed WelfordVar inside the persistent NUTS
    launch) against the same run done iteration by iteration: single-transition launches on the same Philox streams,
    with the ORACLE's vectorised adaptors (stepsize.jl:178-210, massmatrix.jl:141-157) and the reference's window
    schedule / reset / finalize order (stan_adaptor.jl:13-50, 137-159; sampler.jl:72-90) applied on the host."""
    from oracle import oracle_c as oc

    D, N, T, n_adapts = 8, 96, 60, 50
    ib, tb, wsz = 10, 8, 6
    ws, we, splits = oc.stan_windows(n_adapts, ib, tb, wsz)
    assert (ws, we, list(splits)) == (11, 42, [16, 42])  # first window has 6 < n_min draws: reset without update
    rng = np.random.default_rng(3)
    sd = np.exp(rng.uniform(-1.0, 1.0, D))
    target = A.DiagGaussian(rng.normal(size=D), sd)
    th0 = torch.as_tensor(rng.normal(size=(N, D)), device=DEV)
    eps0, seed = 0.3, 99

    h = A.Hamiltonian(A.DiagEuclideanMetric(np.ones(D)), target)
    kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(eps0), A.GeneralisedNoUTurn(8, 1000.0)))
    z0 = A.phasepoint(h, th0, torch.zeros_like(th0))
    adaptor = A.VectorisedStanAdaptor(delta=0.8, init_buffer=ib, term_buffer=tb, window_size=wsz)
    zl, draws, st, eps_f, minv_f, trace = A.nuts_adapt_sample(A.PhiloxRNG(seed), h, kern, z0, T, n_adapts, adaptor,
                                                              keep_eps_trace=True)

    # ---- the same run, one launch per iteration.  The transitions use the step sizes / metric the fused launch
    # reported (so a last-bit difference between device and host libm cannot flip a tree decision and fork the two
    # runs); the ORACLE adaptors run alongside on the loop's acceptance rates and draws and must reproduce the
    # fused launch's next step size at every iteration, its metric at the window end and its final eps.
    prng = A.PhiloxRNG(seed)
    da, wv = oc.DualAveraging(np.full(N, eps0), delta=0.8), oc.WelfordVar((D, N))
    Minv_or, Minv_dev = np.ones((N, D)), torch.ones((N, D), dtype=torch.float64, device=DEV)
    z = z0
    for i in range(1, T + 1):
        assert np.allclose(trace[i - 1].cpu().numpy(), da.eps, rtol=1e-10, atol=0), i
        hi = A.Hamiltonian(A.DiagEuclideanMetric(Minv_dev), target)
        ki = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(trace[i - 1].clone()), A.GeneralisedNoUTurn(8, 1000.0)))
        tr = A.transition(prng, hi, ki, z)
        z = tr.z
        assert rel_err(draws[i - 1].cpu().numpy(), z.theta.cpu().numpy()) < 1e-10, i
        assert (st["n_steps"][i - 1].cpu().numpy() == tr.stat["n_steps"].cpu().numpy()).all(), i
        assert np.allclose(st["acceptance_rate"][i - 1].cpu().numpy(), tr.stat["acceptance_rate"].cpu().numpy(), rtol=1e-10), i
        if i <= n_adapts:
            da.adapt(tr.stat["acceptance_rate"].cpu().numpy())
            if ws <= i <= we:
                wv.push(z.theta.cpu().numpy().T)
                if i in splits and wv.n.value >= 10:
                    Minv_or = np.ascontiguousarray(wv.estimate().T)
                    assert np.allclose(minv_f.cpu().numpy(), Minv_or, rtol=1e-9)  # the only update of this schedule
                    Minv_dev = minv_f
            if i in splits:
                da.reset()
                wv = oc.WelfordVar((D, N))
            if i == n_adapts:
                da.finalize()
    assert np.allclose(eps_f.cpu().numpy(), da.eps, rtol=1e-10)
    assert not np.allclose(Minv_or, 1.0)                    # the second window did update the metric
    assert rel_err(zl.theta.cpu().numpy(), z.theta.cpu().numpy()) < 1e-10
    # and it adapts: per-chain acceptance after warm-up is near delta, M^-1 tracks the target variances
    acc = st["acceptance_rate"][n_adapts:].double().mean().item()
    assert 0.6 < acc < 0.95, acc
    ratio = np.median(minv_f.cpu().numpy(), axis=0) / sd ** 2
    assert 0.4 < ratio.min() and ratio.max() < 2.5  # 26 draws per chain: a rough estimate, as in the reference


def test_sample_with_vectorised_adaptor_runs_funnel_warmup_in_one_launch():
    D, N = 20, 512
    h = A.Hamiltonian(A.DiagEuclideanMetric(np.ones(D)), A.Funnel(D))
    kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.1), A.GeneralisedNoUTurn()))
    th0 = torch.as_tensor(np.random.default_rng(4).normal(size=(N, D)) * 0.1, device=DEV)
    res = ad.sample(A.PhiloxRNG(8), h, kern, th0, 300, A.VectorisedStanAdaptor(delta=0.8), 250)
    acc = np.mean([s["acceptance_rate"] for s in res.stats[250:]])
    assert 0.65 < acc < 0.95, acc
    assert res.eps.shape == (N,) and res.Minv.shape == (N, D)
    assert float(res.eps.min()) > 1e-3 and float(res.eps.max()) < 2.0
    v = res.theta[:, 0].cpu().numpy()
    assert abs(v.std() - 3.0) < 1.0  # the funnel's neck variable ~ N(0, 3^2)


def test_in_launch_adaptation_degenerate_and_host_forms():
    """n_adapts = 0 -> the adaptive family is bit-identical to the plain persistent launch (same arithmetic, same
    Philox streams); host (numpy) buffers give the device result; unsupported combinations fail loudly."""
    D, N, T = 12, 130, 15
    rng = np.random.default_rng(9)
    sd = np.exp(rng.uniform(-0.7, 0.7, D))
    target = A.DiagGaussian(rng.normal(size=D), sd)
    Minv = np.exp(rng.uniform(-0.3, 0.3, D))
    h = A.Hamiltonian(A.DiagEuclideanMetric(Minv), target)
    kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.25), A.GeneralisedNoUTurn(7, 1000.0)))
    th = rng.normal(size=(N, D))
    z0 = A.phasepoint(h, torch.as_tensor(th, device=DEV), torch.zeros((N, D), dtype=torch.float64, device=DEV))
    ad0 = A.VectorisedStanAdaptor()
    zl, dr, st, eps, minv, tr = A.nuts_adapt_sample(A.PhiloxRNG(3), h, kern, z0, T, 0, ad0, keep_eps_trace=True)
    zl2, dr2, st2 = A.sample_transitions(A.PhiloxRNG(3), h, kern, z0, T)
    assert torch.equal(dr, dr2) and torch.equal(zl.theta, zl2.theta) and torch.equal(st["n_steps"], st2["n_steps"])
    assert torch.equal(eps, torch.full_like(eps, 0.25)) and torch.equal(tr, torch.full_like(tr, 0.25))
    assert np.allclose(minv.cpu().numpy(), np.broadcast_to(Minv, (N, D)))
    # adapting run: host buffers == device buffers
    ad1 = A.VectorisedStanAdaptor(init_buffer=3, term_buffer=2, window_size=4, n_min=3)
    zd, dd, sd_, ed, md, _ = A.nuts_adapt_sample(A.PhiloxRNG(4), h, kern, z0, T, 12, ad1)
    zh0 = A.phasepoint(h, th, np.zeros((N, D)))
    zh, dh, sh, eh, mh, _ = A.nuts_adapt_sample(A.PhiloxRNG(4), h, kern, zh0, T, 12, ad1)
    assert np.array_equal(dh, dd.cpu().numpy()) and np.array_equal(eh, ed.cpu().numpy()) and np.array_equal(mh, md.cpu().numpy())
    assert not np.allclose(mh, np.broadcast_to(Minv, (N, D)))  # the windows did update the metric
    assert (eh != 0.25).all()
    # per-chain initial step sizes are honoured
    e0 = torch.as_tensor(rng.uniform(0.1, 0.4, N), device=DEV)
    k2 = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(e0), A.GeneralisedNoUTurn(7, 1000.0)))
    _, _, _, e2, _, t2 = A.nuts_adapt_sample(A.PhiloxRNG(5), h, k2, z0, 3, 0, ad0, keep_eps_trace=True)
    assert torch.equal(t2[0], e0) and torch.equal(e2, e0)
    # loud failures
    hu = A.Hamiltonian(A.UnitEuclideanMetric(D), target)
    with pytest.raises(A.AhmcError):
        A.nuts_adapt_sample(A.PhiloxRNG(1), hu, kern, A.phasepoint(hu, z0.theta, z0.r), 4, 2, ad0)
    ks = A.HMCKernel(A.Trajectory(A.SliceTS, A.Leapfrog(0.25), A.GeneralisedNoUTurn()))
    with pytest.raises(A.AhmcError):
        A.nuts_adapt_sample(A.PhiloxRNG(1), h, ks, z0, 4, 2, ad0)
    with pytest.raises(A.AhmcError):
        A.nuts_adapt_sample(A.PhiloxRNG(1), h, kern, z0, 4, 5, ad0)  # n_adapts > n_transitions


@pytest.mark.gpu
def test_in_launch_adaptation_on_a_dense_precision_target_runs_the_cooperative_form():
    """A dense-precision Gaussian with the Diag metric runs the adaptive family in the block-cooperative form (8 chains
    share the precision product).  Written with a DIAGONAL precision it is the diagonal Gaussian: same trees, and step
    sizes / adapted M^-1 / draws equal up to the summation order of the products."""
    D, N, T, n_adapts = 64, 50, 14, 12
    rng = np.random.default_rng(19)
    sd, mu = np.exp(rng.uniform(-0.4, 0.4, D)), rng.normal(size=D) * 0.3
    kern = A.HMCKernel(A.Trajectory(A.MultinomialTS, A.Leapfrog(0.2), A.GeneralisedNoUTurn(7, 1000.0)))
    ad = A.VectorisedStanAdaptor(init_buffer=3, term_buffer=2, window_size=4, n_min=3)
    th = torch.as_tensor(rng.normal(size=(N, D)), device=DEV)
    out = []
    for target in (A.DiagGaussian(mu, sd), A.DenseGaussian(mu, np.diag(1.0 / sd ** 2))):
        h = A.Hamiltonian(A.DiagEuclideanMetric(np.ones(D)), target)
        z0 = A.phasepoint(h, th, torch.zeros_like(th))
        out.append(A.nuts_adapt_sample(A.PhiloxRNG(8), h, kern, z0, T, n_adapts, ad))
    (zl0, d0, s0, e0, m0, _), (zl1, d1, s1, e1, m1, _) = out
    # the unnormalised dense target has another constant c0: energies differ by it, trees must not
    assert torch.equal(s0["n_steps"], s1["n_steps"]) and torch.equal(s0["tree_depth"], s1["tree_depth"])
    assert torch.allclose(e0, e1, rtol=1e-6) and torch.allclose(m0, m1, rtol=1e-5) and torch.allclose(d0, d1, atol=1e-5)
    assert (e0 != 0.2).all() and not torch.allclose(m0, torch.ones_like(m0))
