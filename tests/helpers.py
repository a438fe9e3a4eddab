"""Shared test helpers: golden-case loading and synthetic inputs (numpy PCG64 seeds)."""
import json
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
MODEL_KINDS = {"std_normal": 0, "diag_gauss": 1, "dense_gauss": 2, "funnel": 3}
METRIC_KINDS = {"unit": 0, "diag": 1, "dense": 2}


def _with_reference_cases(cases, ref_file):
    """append the cases of a reference-generated file (tests/golden/gen_from_reference.jl run against the real
    AdvancedHMC.jl) when it has been committed; their names get the prefix "ref:" """
    path = os.path.join(HERE, "golden", ref_file)
    if os.path.exists(path):
        with open(path) as f:
            for c in json.load(f)["cases"]:
                c = dict(c)
                c["name"] = "ref:" + c["name"]
                cases.append(c)
    return cases


def reference_fixtures_present():
    return all(os.path.exists(os.path.join(HERE, "golden", f)) for f in ("leapfrog_ref.json", "hmc_ref.json", "nuts_ref.json"))


def golden_cases():
    with open(os.path.join(HERE, "golden", "leapfrog_mp50.json")) as f:
        d = json.load(f)
    d["cases"] = _with_reference_cases(d["cases"], "leapfrog_ref.json")
    return d


def nuts_golden_cases():
    """tests/golden/nuts_mp50.json: NUTS transitions from the recursive 50-digit restatement (gen_nuts_mp.py)."""
    with open(os.path.join(HERE, "golden", "nuts_mp50.json")) as f:
        return _with_reference_cases(json.load(f)["cases"], "nuts_ref.json")


def hmc_golden_cases():
    """tests/golden/hmc_mp50.json: static HMC transitions (refresh, EndPointTS / MultinomialTS) from gen_hmc_mp.py."""
    with open(os.path.join(HERE, "golden", "hmc_mp50.json")) as f:
        return _with_reference_cases(json.load(f)["cases"], "hmc_ref.json")


def case_arrays(case):
    """-> dict of numpy arrays in (D,N) Fortran layout."""
    D, N = case["D"], case["N"]
    f = lambda rows: np.asfortranarray(np.array(rows, dtype=np.float64).T)
    out = dict(theta0=f(case["theta0"]), r0=f(case["r0"]))
    e = case["expect"]
    out["theta"], out["r"], out["lp_gradient"] = f(e["theta"]), f(e["r"]), f(e["lp_gradient"])
    out["lp_value"], out["lk_value"] = np.array(e["lp_value"]), np.array(e["lk_value"])
    out["p0"] = None if case["p0"] is None else np.array(case["p0"], dtype=np.float64)
    p1 = case["p1"]
    out["p1"] = None if p1 is None else np.asfortranarray(np.array(p1, dtype=np.float64))
    Mi = case["Minv"]
    out["Minv"] = None if Mi is None else np.asfortranarray(np.array(Mi, dtype=np.float64))
    out["eps"] = case["eps"] if case["eps_chain"] is None else np.array(case["eps_chain"])
    assert out["theta0"].shape == (D, N)
    return out


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny) -- the 1e-10 relative fp64 criterion of BASELINE.json."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    scale = max(float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b))) / scale


def rel_err_elem(a, b, floor=1e-6):
    """element-wise relative error with an absolute floor (for values that pass through zero)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def rel_err_elem_scaled(a, b, floor_frac=1e-3):
    """ELEMENT-WISE relative error, max_i |a_i - b_i| / max(|b_i|, floor_frac * max|b|): every coordinate must carry
    its own correct digits; only coordinates below floor_frac of the largest one (values passing through zero) are
    measured against that floor instead of against themselves.  The parity criterion of the state comparisons
    (DESIGN.md section 4); `rel_err` (max-norm relative) is kept beside it."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    fin = np.isfinite(b)
    if not fin.all():  # non-finite reference entries must match exactly; the rest is compared numerically
        if not np.array_equal(a[~fin], b[~fin], equal_nan=True):
            return float("inf")
        a, b = a[fin], b[fin]
        if a.size == 0:
            return 0.0
    floor = max(floor_frac * float(np.max(np.abs(b))), 1e-300)
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def synth_diag_gauss(D, N, seed, s_lo=0.1, s_hi=10.0):
    """North-star headline shape: diagonal Gaussian, s log-spaced, Minv = s^2,
    theta0 ~ N(0,1), r0 ~ N(0, M) i.e. z / sqrt(Minv)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    s = np.exp(np.linspace(np.log(s_lo), np.log(s_hi), D))
    m = np.zeros(D)
    Minv = s * s
    theta = np.asfortranarray(rng.normal(size=(N, D)).T)
    r = np.asfortranarray((rng.normal(size=(N, D)) / np.sqrt(Minv)).T)
    return m, s, Minv, theta, r
