"""CPU test of the per-chain NUTS state machine prototyped for the tile-centric kernel
(advancedhmc.jl_b200/csrc/experimental/ahmc_tree_logic.cuh): a g++-built harness plays the vector half, the recursive C
oracle is the checker.  No GPU involved; the header is not part of the shipped library yet."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle_c as oc
from tests.helpers import rel_err

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    out = tmp_path_factory.mktemp("tree_logic") / "libtree_logic.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off",
                    "-I", os.path.join(ROOT, "advancedhmc.jl_b200", "csrc", "experimental"),
                    os.path.join(ROOT, "tests", "tree_logic_harness.cpp"), "-o", str(out)], check=True)
    return ctypes.CDLL(str(out))


@pytest.mark.parametrize("eps,delta_max", [(0.12, 1000.0), (0.4, 0.3)], ids=["deep", "divergent"])
@pytest.mark.parametrize("sampler,criterion", [("multinomial", "generalised"), ("slice", "generalised"),
                                               ("multinomial", "classic"), ("multinomial", "strict"),
                                               ("slice", "classic"), ("slice", "strict")])
def test_state_machine_reproduces_the_recursive_oracle(harness, sampler, criterion, eps, delta_max):
    rng = np.random.default_rng(42)
    D, N, max_depth = 6, 150, 8
    mu, sd, Minv = rng.normal(size=D), np.exp(rng.uniform(-1, 1, D)), np.exp(rng.uniform(-0.5, 0.5, D))
    th, r = rng.normal(size=(N, D)) * 2.0, rng.normal(size=(N, D))
    dirs = rng.integers(0, 2, size=(N, max_depth + 1)).astype(np.uint8)
    var = rng.exponential(size=(N, 1 << max_depth))
    if sampler == "slice":
        var[:, 1:] = rng.uniform(size=(N, (1 << max_depth) - 1))
    model, metric = oc.Model(oc.DIAG_GAUSS, D, mu, sd, 0.0), oc.Metric(oc.DIAG, Minv)
    zo, so, used = oc.nuts_transition(model, metric, eps, oc.phasepoint(model, metric, th.T, r.T), None, dirs, var,
                                      max_depth=max_depth, delta_max=delta_max, sampler=sampler, criterion=criterion)
    tho, ro = np.zeros((N, D)), np.zeros((N, D))
    ns, dp, ne, vu = (np.zeros(N, dtype=np.int32) for _ in range(4))
    ar, dh = np.zeros(N), np.zeros(N)
    P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    rc = harness.tree_logic_nuts(D, ctypes.c_int64(N), P(mu), P(sd), P(Minv), ctypes.c_double(eps), oc.SAMPLER[sampler],
                                 oc.CRITERION[criterion], max_depth, ctypes.c_double(delta_max), P(th), P(r), P(dirs),
                                 ctypes.c_int64(dirs.shape[1]), P(var), ctypes.c_int64(var.shape[1]), P(tho), P(ro), P(ns),
                                 P(dp), P(ne), P(ar), P(dh), P(vu))
    assert rc == 0
    assert (dp == so.tree_depth).all() and (ns == so.n_steps).all() and (ne == so.numerical_error).all()
    assert (vu == used).all()
    assert rel_err(tho.T, zo.theta) < 1e-10 and rel_err(ro.T, zo.r) < 1e-10
    assert np.allclose(ar, so.acceptance_rate, rtol=1e-10)
    assert np.allclose(dh, so.max_hamiltonian_energy_error, rtol=1e-9, atol=1e-12)
    if delta_max < 1.0:
        assert sampler == "slice" or so.numerical_error.sum() > 5  # (SliceTS: Delta_max is measured from the slice level)
    else:
        assert len(set(so.tree_depth)) >= 3 and (((so.n_steps + 1) & so.n_steps) != 0).any()
