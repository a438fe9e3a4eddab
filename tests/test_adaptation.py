"""CPU tests of the host-side pooled adaptation logic (advancedhmc.jl_b200/adaptation.py) against the oracle's
restatement of src/adaptation/*.jl, plus the world_size-2 gloo path of the adaptor-record exchange."""
import os
import socket

import numpy as np
import pytest

import ahmc_b200 as A
from ahmc_b200 import adaptation as ad
from oracle import oracle_c as oc


def _record(theta, alpha):
    """what ahmc_adapt_summary_f64 produces for one rank: [N, sum min(1,a), mean, M2]."""
    N, D = theta.shape
    mu = theta.mean(axis=0)
    return np.concatenate([[N, np.minimum(1, alpha).sum()], mu, ((theta - mu) ** 2).sum(axis=0)])


def test_merge_records_equals_pooled_statistics():
    rng = np.random.default_rng(0)
    parts = [rng.normal(size=(n, 7)) * 2 + 1 for n in (5, 1, 64, 3)]
    alphas = [rng.uniform(0, 1.5, size=p.shape[0]) for p in parts]
    merged = ad.merge_records([_record(p, a) for p, a in zip(parts, alphas)])
    allx, alla = np.concatenate(parts), np.concatenate(alphas)
    want = _record(allx, alla)
    assert np.allclose(merged, want, rtol=1e-12, atol=1e-12)


def test_dual_averaging_one_chain_equals_reference_path():
    """pooled DA with a single chain == stepsize.jl:178-210 (oracle)."""
    rng = np.random.default_rng(1)
    da, ref = ad.NesterovDualAveraging(0.8, 0.1), oc.DualAveraging([0.1], delta=0.8)
    for _ in range(60):
        a = rng.uniform(0, 1.4)
        da.adapt(min(1.0, a))
        ref.adapt([a])
        assert da.eps == pytest.approx(ref.eps[0], rel=1e-14) and da.m == ref.m
    da.reset(), ref.reset()
    assert da.mu == pytest.approx(ref.mu[0], rel=1e-15) and da.m == 0
    for _ in range(5):
        da.adapt(0.5), ref.adapt([0.5])
    da.finalize(), ref.finalize()
    assert da.eps == pytest.approx(ref.eps[0], rel=1e-14)


def test_pooled_welford_one_chain_equals_reference_and_many_chains_equal_pushing_each():
    rng = np.random.default_rng(2)
    D = 6
    w, ref = ad.WelfordVar(D), oc.WelfordVar((D,))
    for _ in range(30):  # one chain per iteration: Chan merge of (1, x, 0) == massmatrix.jl:141-149
        x = rng.normal(size=D)
        w.push_record(_record(x[None, :], np.ones(1)))
        ref.push(x)
    assert np.allclose(w.mu, ref.mu, rtol=1e-13) and np.allclose(w.M, ref.M, rtol=1e-12)
    assert np.allclose(w.get_estimation(), ref.estimate(), rtol=1e-12)
    # 16 chains per iteration == pushing the 16 positions one after another into the reference estimator
    w2, ref2 = ad.WelfordVar(D), oc.WelfordVar((D,))
    for _ in range(10):
        X = rng.normal(size=(16, D)) * np.arange(1, D + 1)
        w2.push_record(_record(X, np.ones(16)))
        for x in X:
            ref2.push(x)
    assert w2.n == 160 and np.allclose(w2.get_estimation(), ref2.estimate(), rtol=1e-11)


def test_stan_windows_pin_and_adaptor_schedule():
    """test/adaptation.jl:131-151."""
    assert ad.stan_windows(1000) == (76, 950, [100, 150, 250, 450, 950])
    assert ad.stan_windows(1000) == oc.stan_windows(1000)
    for n in (100, 150, 200, 537, 2000):
        assert ad.stan_windows(n) == oc.stan_windows(n)
    D = 3
    rng = np.random.default_rng(3)
    adp = ad.StanHMCAdaptor(ad.WelfordVar(D), ad.NesterovDualAveraging(0.8, 0.1))
    adp.initialize(1000)
    updates = []
    for i in range(1, 1001):
        before = adp.pc.var.copy()
        adp.adapt(_record(rng.normal(size=(8, D)) * [1, 2, 3], rng.uniform(0.5, 1, 8)))
        if not np.array_equal(before, adp.pc.var):
            updates.append(i)
        if i in (100, 150, 250, 450, 950):
            assert adp.pc.n == 0 and adp.ssa.m == 0  # reset at window ends (stan_adaptor.jl:155-158)
    adp.finalize()
    assert updates == [100, 150, 250, 450, 950]
    assert np.allclose(adp.Minv, [1, 4, 9], rtol=0.15)
    assert adp.eps == pytest.approx(np.exp(adp.ssa.x_bar))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(100 + rank)
    theta, alpha = rng.normal(size=(5 + rank, 4)), rng.uniform(0, 1.2, 5 + rank)
    rec = torch.as_tensor(_record(theta, alpha))
    merged = ad.merge_records(ad.allgather_records(rec))
    # the dense and the position+gradient record layouts go through the same single all-gather
    grad = -theta * 0.5
    rec_cov = torch.as_tensor(_record_cov(theta, alpha))
    rec_nut = torch.as_tensor(_record_nutpie(theta, grad, alpha))
    merged_cov = ad.merge_records(ad.allgather_records(rec_cov), "cov")
    merged_nut = ad.merge_records(ad.allgather_records(rec_nut), "nutpie")
    # a short pooled Stan warm-up driven by exchanged records: every rank must end with the same eps and M^-1
    adaptor = ad.StanHMCAdaptor(ad.WelfordVar(4), ad.NesterovDualAveraging(0.8, 0.1), init_buffer=3, term_buffer=2, window_size=4)
    adaptor.initialize(20)
    for it in range(20):
        th_i = rng.normal(size=(6 + rank, 4)) * (1.0 + 0.1 * it)
        al_i = rng.uniform(0.3, 1.1, 6 + rank)
        adaptor.adapt(ad.merge_records(ad.allgather_records(torch.as_tensor(_record(th_i, al_i)))))
    adaptor.finalize()
    q.put((rank, merged, merged_cov, merged_nut, adaptor.eps, np.array(adaptor.Minv)))
    dist.destroy_process_group()


def test_adaptor_record_allgather_world_size_2_gloo():
    """The path's only exchange: every rank ends with the same, rank-ordered merge."""
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    got = [q.get(timeout=120) for _ in range(2)]
    [p.join(timeout=60) for p in ps]
    full = {g[0]: g[1:] for g in got}
    res = {r: full[r][0] for r in full}
    assert np.array_equal(res[0], res[1])  # bit-identical on every rank
    for k in (1, 2, 4):  # cov record, nutpie record, adapted M^-1: bit-identical across ranks too
        assert np.array_equal(full[0][k], full[1][k])
    assert full[0][3] == full[1][3] and 0.0 < full[0][3] < 10.0  # the pooled step size
    assert not np.allclose(full[0][4], 1.0)  # the window did update the metric
    parts = [(np.random.default_rng(100 + r).normal(size=(5 + r, 4)), None) for r in range(2)]
    th = [np.random.default_rng(100 + r) for r in range(2)]
    recs = []
    for r in range(2):
        g = np.random.default_rng(100 + r)
        t, a = g.normal(size=(5 + r, 4)), g.uniform(0, 1.2, 5 + r)
        recs.append(_record(t, a))
    assert np.array_equal(res[0], ad.merge_records(recs))
    assert res[0][0] == 11
    allx = np.concatenate([np.random.default_rng(100 + r).normal(size=(5 + r, 4)) for r in range(2)])
    c = allx - allx.mean(axis=0)
    assert np.allclose(full[0][1][2 + 8:].reshape(4, 4), c.T @ c, rtol=1e-12, atol=1e-12)


def _welford_cov_reference(xs):
    """sequential restatement of `push!(::WelfordCov, s)` + `get_estimation` (src/adaptation/massmatrix.jl:324-340)."""
    D = xs.shape[1]
    n, mu, M = 0, np.zeros(D), np.zeros((D, D))
    for s in xs:
        n += 1
        delta = s - mu
        mu = mu + delta / n
        M = M + np.outer(s - mu, delta)
    return n, mu, M, n / ((n + 5) * (n - 1)) * M + 1e-3 * (5 / (n + 5)) * np.eye(D)


def _record_cov(theta, alpha):
    mu = theta.mean(axis=0)
    c = theta - mu
    return np.concatenate([_record(theta, alpha), (c.T @ c).ravel()])


def _record_nutpie(theta, grad, alpha):
    return np.concatenate([_record(theta, alpha), _record(grad, alpha)[2:]])


def test_merge_records_cov_and_nutpie_layouts():
    rng = np.random.default_rng(3)
    D = 5
    L = rng.normal(size=(D, D))
    parts = [rng.normal(size=(n, D)) @ L + 0.5 for n in (4, 1, 33, 2)]
    grads = [-(p - 0.5) @ np.linalg.inv(L @ L.T) for p in parts]
    alphas = [rng.uniform(0, 1.5, size=p.shape[0]) for p in parts]
    allx, allg, alla = np.concatenate(parts), np.concatenate(grads), np.concatenate(alphas)
    got = ad.merge_records([_record_cov(p, a) for p, a in zip(parts, alphas)], "cov")
    assert np.allclose(got, _record_cov(allx, alla), rtol=1e-12, atol=1e-12)
    got = ad.merge_records([_record_nutpie(p, g, a) for p, g, a in zip(parts, grads, alphas)], "nutpie")
    assert np.allclose(got, _record_nutpie(allx, allg, alla), rtol=1e-12, atol=1e-12)


def test_pooled_welford_cov_and_nutpie_equal_the_sequential_reference_estimators():
    rng = np.random.default_rng(4)
    D = 4
    L = rng.normal(size=(D, D))
    xs = rng.normal(size=(48, D)) @ L
    gs = -xs @ np.linalg.inv(L @ L.T)
    n, mu, M, est = _welford_cov_reference(xs)
    # one chain per record (the reference's own granularity) and 16 chains per record give the same estimator
    for chunk in (1, 16):
        wc, nv = ad.WelfordCov(D), ad.NutpieVar(D)
        for k in range(0, len(xs), chunk):
            a = np.ones(chunk)
            wc.push_record(_record_cov(xs[k:k + chunk], a))
            nv.push_record(_record_nutpie(xs[k:k + chunk], gs[k:k + chunk], a))
        assert wc.n == n and np.allclose(wc.mu, mu, rtol=1e-12) and np.allclose(wc.M, M, rtol=1e-11, atol=1e-12)
        assert np.allclose(wc.get_estimation(), est, rtol=1e-11, atol=1e-13)
        wv_x, wv_g = oc.WelfordVar((D,)), oc.WelfordVar((D,))
        for x, g in zip(xs, gs):
            wv_x.push(x), wv_g.push(g)
        assert np.allclose(nv.get_estimation(), np.sqrt(wv_x.estimate() / wv_g.estimate()), rtol=1e-11)
    wc.update(), nv.update()
    assert wc.var.shape == (D, D) and nv.var.shape == (D,)
    # for a Gaussian the Nutpie estimate is sqrt(var_x * var of (Sigma^-1 x))^-1 ... -> positive, finite
    assert np.all(np.isfinite(nv.var)) and np.all(nv.var > 0)


def test_adaptors_match_mp50_known_answers():
    """oracle AND host-mirror adaptors against tests/golden/adapt_mp50.json (50-digit restatement of stepsize.jl:25-62,
    178-210 and massmatrix.jl:141-157, 244-248, 324-340; tests/golden/gen_adapt_mp.py)."""
    import json

    with open(os.path.join(os.path.dirname(__file__), "golden", "adapt_mp50.json")) as f:
        g = json.load(f)
    d = g["dual_averaging"]
    da_o, da_h = oc.DualAveraging([d["eps0"]], delta=d["delta"]), ad.NesterovDualAveraging(d["delta"], d["eps0"])
    for i, (a, want) in enumerate(zip(d["alphas"], d["eps_trace"]), 1):
        da_o.adapt([a])
        da_h.adapt(min(1.0, a))
        assert da_o.eps[0] == pytest.approx(want, rel=1e-12) and da_h.eps == pytest.approx(want, rel=1e-12)
        if i == d["reset_at"]:
            da_o.reset(), da_h.reset()
    da_o.finalize(), da_h.finalize()
    assert da_o.eps[0] == pytest.approx(d["eps_final"], rel=1e-12) and da_h.eps == pytest.approx(d["eps_final"], rel=1e-12)
    w = g["welford"]
    xs, gs = np.array(w["xs"]), np.array(w["gs"])
    D = xs.shape[1]
    wo, co = oc.WelfordVar((D,)), oc.WelfordCov(D)
    wh, ch, nh = ad.WelfordVar(D), ad.WelfordCov(D), ad.NutpieVar(D)
    for x, gr in zip(xs, gs):
        wo.push(x), co.push(x)
        one = np.ones(1)
        wh.push_record(_record(x[None, :], one))
        ch.push_record(_record_cov(x[None, :], one))
        nh.push_record(_record_nutpie(x[None, :], gr[None, :], one))
    for est in (wo.estimate(), wh.get_estimation()):
        assert np.allclose(est, w["var_estimate"], rtol=1e-11)
    for est in (co.estimate(), ch.get_estimation()):
        assert np.allclose(est, w["cov_estimate"], rtol=1e-10, atol=1e-13)
    assert np.allclose(wh.mu, w["mu"], rtol=1e-12) and np.allclose(wh.M, w["M"], rtol=1e-11)
    assert np.allclose(nh.get_estimation(), w["nutpie_estimate"], rtol=1e-10)
