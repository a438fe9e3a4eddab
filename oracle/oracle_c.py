"""oracle_c.py -- ctypes binding of oracle/_build/libahmc_oracle.so (the C restatement).

*** TEST / BENCH INFRASTRUCTURE ONLY: never imported by the product package. ***
Arrays are float64, shape (D, N), Fortran order (Julia column-major D x N).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libahmc_oracle.so")

STD_NORMAL, DIAG_GAUSS, DENSE_GAUSS, FUNNEL = 0, 1, 2, 3
UNIT, DIAG, DENSE = 0, 1, 2

_dp = C.POINTER(C.c_double)


class _Model(C.Structure):
    _fields_ = [("kind", C.c_int32), ("D", C.c_int32), ("p0", _dp), ("p1", _dp), ("c0", C.c_double)]


class _Metric(C.Structure):
    _fields_ = [("kind", C.c_int32), ("Minv", _dp), ("chain_stride", C.c_int64), ("cholU", _dp)]


class _PP(C.Structure):
    _fields_ = [("theta", _dp), ("r", _dp), ("lp_value", _dp), ("lp_gradient", _dp), ("lk_value", _dp),
                ("lk_gradient", _dp), ("ld", C.c_int64)]


class _Stats(C.Structure):
    _fields_ = [("n_steps", C.POINTER(C.c_int32)), ("is_accept", C.POINTER(C.c_uint8)),
                ("acceptance_rate", _dp), ("log_density", _dp), ("hamiltonian_energy", _dp),
                ("hamiltonian_energy_error", _dp), ("max_hamiltonian_energy_error", _dp),
                ("tree_depth", C.POINTER(C.c_int32)), ("numerical_error", C.POINTER(C.c_uint8))]


class _DA(C.Structure):
    _fields_ = [("m", C.c_int64), ("eps", _dp), ("mu", _dp), ("x_bar", _dp), ("H_bar", _dp)]


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (recipe: oracle/Makefile). Building the checker is not using it."""
    if force or not os.path.exists(_LIB_PATH):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.orc_neg_kinetic.restype = C.c_double
        _lib.orc_stan_windows.restype = C.c_int32
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _f(a):
    return np.asfortranarray(np.asarray(a, dtype=np.float64))


class Model:
    def __init__(self, kind, D, p0=None, p1=None, c0=0.0):
        self.kind, self.D, self.c0 = kind, int(D), float(c0)
        self.p0 = None if p0 is None else _f(p0)
        self.p1 = None if p1 is None else _f(p1)
        self.c = _Model(kind, self.D, _p(self.p0), _p(self.p1), self.c0)


class Metric:
    def __init__(self, kind, Minv=None):
        self.kind = kind
        self.Minv = None if Minv is None else _f(Minv)
        self.cholU = None
        stride = 0
        if kind == DIAG and self.Minv.ndim == 2:
            stride = self.Minv.shape[0]
        if kind == DENSE:
            self.cholU = _f(np.linalg.cholesky(self.Minv).T)
        self.c = _Metric(kind, _p(self.Minv), stride, _p(self.cholU))


class PhasePoint:
    """SoA phase point; arrays (D,N) Fortran float64 + (N,) vectors."""

    def __init__(self, D, N, theta=None, r=None, with_lk_gradient=True):
        self.D, self.N = D, N
        z = lambda: np.zeros((D, N), order="F")
        self.theta = z() if theta is None else _f(theta).copy(order="F")
        self.r = z() if r is None else _f(r).copy(order="F")
        self.lp_value = np.zeros(N)
        self.lp_gradient = z()
        self.lk_value = np.zeros(N)
        self.lk_gradient = z() if with_lk_gradient else None

    @property
    def c(self):
        return _PP(_p(self.theta), _p(self.r), _p(self.lp_value), _p(self.lp_gradient), _p(self.lk_value),
                   _p(self.lk_gradient), self.D)

    def energy(self):
        return -(self.lp_value + self.lk_value)


class Stats:
    def __init__(self, N):
        self.n_steps = np.zeros(N, dtype=np.int32)
        self.is_accept = np.zeros(N, dtype=np.uint8)
        self.acceptance_rate = np.zeros(N)
        self.log_density = np.zeros(N)
        self.hamiltonian_energy = np.zeros(N)
        self.hamiltonian_energy_error = np.zeros(N)
        self.max_hamiltonian_energy_error = np.zeros(N)
        self.tree_depth = np.zeros(N, dtype=np.int32)
        self.numerical_error = np.zeros(N, dtype=np.uint8)

    @property
    def c(self):
        i32 = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        u8 = lambda a: a.ctypes.data_as(C.POINTER(C.c_uint8))
        return _Stats(i32(self.n_steps), u8(self.is_accept), _p(self.acceptance_rate), _p(self.log_density),
                      _p(self.hamiltonian_energy), _p(self.hamiltonian_energy_error),
                      _p(self.max_hamiltonian_energy_error), i32(self.tree_depth), u8(self.numerical_error))


def dHdr(metric, r, c=0):
    """dH/dr of one chain (hamiltonian.jl:50-68)."""
    r = np.ascontiguousarray(r, dtype=np.float64)
    out = np.zeros_like(r)
    lib().orc_dHdr(C.byref(metric.c), C.c_int32(r.size), C.c_int64(c), _p(r), _p(out))
    return out


def neg_kinetic(metric, r, c=0):
    """kinetic part of neg_energy for one chain (hamiltonian.jl:155-184)."""
    r = np.ascontiguousarray(r, dtype=np.float64)
    return float(lib().orc_neg_kinetic(C.byref(metric.c), C.c_int32(r.size), C.c_int64(c), _p(r)))


def phasepoint(model, metric, theta, r):
    D, N = theta.shape
    z = PhasePoint(D, N, theta, r)
    zc = z.c
    lib().orc_make_phasepoint(C.byref(model.c), C.byref(metric.c), D, C.c_int64(N), C.byref(zc))
    return z


def _eps_args(eps, N):
    if np.ndim(eps) == 0:
        return C.c_double(float(eps)), None, None
    e = np.ascontiguousarray(eps, dtype=np.float64)
    assert e.shape == (N,)
    return C.c_double(0.0), _p(e), e


def leapfrog(model, metric, eps, z, n_steps, temper_alpha=0.0, compat_break_all=False):
    D, N = z.theta.shape
    out = PhasePoint(D, N)
    status = np.zeros(N, dtype=np.uint32)
    done = np.zeros(N, dtype=np.int32)
    e, ep, _keep = _eps_args(eps, N)
    zc, oc = z.c, out.c
    lib().orc_leapfrog(C.byref(model.c), C.byref(metric.c), D, C.c_int64(N), e, ep, int(n_steps),
                       C.c_double(temper_alpha), C.byref(zc), C.byref(oc),
                       status.ctypes.data_as(C.POINTER(C.c_uint32)), done.ctypes.data_as(C.POINTER(C.c_int32)),
                       int(compat_break_all))
    return out, status, done


def leapfrog_trajectory(model, metric, eps, z, n_steps, temper_alpha=0.0, compat_break_all=False):
    D, N = z.theta.shape
    L = abs(n_steps)
    traj = dict(theta=np.zeros((D, N, L), order="F"), r=np.zeros((D, N, L), order="F"),
                lp_gradient=np.zeros((D, N, L), order="F"), lk_gradient=np.zeros((D, N, L), order="F"),
                lp_value=np.zeros((N, L), order="F"), lk_value=np.zeros((N, L), order="F"))
    tc = _PP(_p(traj["theta"]), _p(traj["r"]), _p(traj["lp_value"]), _p(traj["lp_gradient"]),
             _p(traj["lk_value"]), _p(traj["lk_gradient"]), D)
    done = np.zeros(N, dtype=np.int32)
    e, ep, _keep = _eps_args(eps, N)
    zc = z.c
    lib().orc_leapfrog_trajectory(C.byref(model.c), C.byref(metric.c), D, C.c_int64(N), e, ep, int(n_steps),
                                  C.c_double(temper_alpha), C.byref(zc), C.byref(tc), C.c_int64(D * N),
                                  done.ctypes.data_as(C.POINTER(C.c_int32)), int(compat_break_all))
    return traj, done


def leapfrog_omp(model, metric, eps, z, n_steps, n_threads=0, out=None):
    D, N = z.theta.shape
    out = out or PhasePoint(D, N, with_lk_gradient=False)
    e, ep, _keep = _eps_args(eps, N)
    zc, oc = z.c, out.c
    lib().orc_leapfrog_omp(C.byref(model.c), C.byref(metric.c), D, C.c_int64(N), e, ep, int(n_steps),
                           C.byref(zc), C.byref(oc), int(n_threads))
    return out


def hmc_transition(model, metric, eps, n_steps, z, normal_tape, exp_tape, compat_break_all=False):
    D, N = z.theta.shape
    out, st = PhasePoint(D, N), Stats(N)
    nt = None if normal_tape is None else _f(normal_tape)
    et = np.ascontiguousarray(exp_tape, dtype=np.float64)
    e, ep, _keep = _eps_args(eps, N)
    zc, oc, sc = z.c, out.c, st.c
    lib().orc_hmc_transition(C.byref(model.c), C.byref(metric.c), D, C.c_int64(N), e, ep, int(n_steps), _p(nt),
                             _p(et), C.byref(zc), C.byref(oc), C.byref(sc), int(compat_break_all))
    return out, st


def hmc_multinomial_transition(model, metric, eps, n_steps, n_steps_fwd, z, normal_tape, unif_tape):
    D, N = z.theta.shape
    out, st = PhasePoint(D, N), Stats(N)
    nt = None if normal_tape is None else _f(normal_tape)
    ut = np.ascontiguousarray(unif_tape, dtype=np.float64)
    e, ep, _keep = _eps_args(eps, N)
    zc, oc, sc = z.c, out.c, st.c
    lib().orc_hmc_multinomial_transition(C.byref(model.c), C.byref(metric.c), D, C.c_int64(N), e, ep, int(n_steps),
                                         int(n_steps_fwd), _p(nt), _p(ut), C.byref(zc), C.byref(oc), C.byref(sc))
    return out, st


SAMPLER = {"multinomial": 0, "slice": 1}
CRITERION = {"generalised": 0, "classic": 1, "strict": 2}


def nuts_transition(model, metric, eps, z, normal_tape, dir_tape, exp_tape, max_depth=10, delta_max=1000.0,
                    sampler="multinomial", criterion="generalised"):
    """dir_tape: (N, n_dir) uint8 C-order; exp_tape: (N, n_exp) float64 C-order."""
    D, N = z.theta.shape
    out, st = PhasePoint(D, N), Stats(N)
    nt = None if normal_tape is None else _f(normal_tape)
    dt = np.ascontiguousarray(dir_tape, dtype=np.uint8)
    et = np.ascontiguousarray(exp_tape, dtype=np.float64)
    used = np.zeros(N, dtype=np.int32)
    e, ep, _keep = _eps_args(eps, N)
    zc, oc, sc = z.c, out.c, st.c
    lib().orc_nuts_transition_ex(C.byref(model.c), C.byref(metric.c), D, C.c_int64(N), e, ep, int(max_depth),
                              C.c_double(delta_max), SAMPLER[sampler], CRITERION[criterion], _p(nt),
                              dt.ctypes.data_as(C.POINTER(C.c_uint8)),
                              C.c_int64(dt.shape[1]), _p(et), C.c_int64(et.shape[1]), C.byref(zc), C.byref(oc),
                              C.byref(sc), used.ctypes.data_as(C.POINTER(C.c_int32)))
    return out, st, used


class DualAveraging:
    """NesterovDualAveraging (stepsize.jl:111-229) over n entries."""

    def __init__(self, eps, delta=0.8, gamma=0.05, t0=10.0, kappa=0.75):
        self.eps = np.atleast_1d(np.asarray(eps, dtype=np.float64)).copy()
        n = self.n = self.eps.size
        self.mu, self.x_bar, self.H_bar = np.zeros(n), np.zeros(n), np.zeros(n)
        self.delta, self.gamma, self.t0, self.kappa = delta, gamma, t0, kappa
        self.s = _DA(0, _p(self.eps), _p(self.mu), _p(self.x_bar), _p(self.H_bar))
        lib().orc_da_init(C.byref(self.s), C.c_int64(n))

    @property
    def m(self):
        return self.s.m

    def adapt(self, alpha):
        a = np.atleast_1d(np.asarray(alpha, dtype=np.float64))
        lib().orc_da_adapt(C.byref(self.s), C.c_int64(self.n), C.c_double(self.gamma), C.c_double(self.t0),
                           C.c_double(self.kappa), C.c_double(self.delta), _p(a))

    def reset(self):
        lib().orc_da_reset(C.byref(self.s), C.c_int64(self.n))

    def finalize(self):
        lib().orc_da_finalize(C.byref(self.s), C.c_int64(self.n))


class WelfordVar:
    def __init__(self, shape):
        self.n = C.c_int64(0)
        self.mu = np.zeros(shape, order="F")
        self.M = np.zeros(shape, order="F")

    def push(self, s):
        s = _f(s)
        lib().orc_welford_var_push(C.byref(self.n), _p(self.mu), _p(self.M), C.c_int64(self.mu.size), _p(s))

    def estimate(self):
        out = np.zeros_like(self.mu, order="F")
        lib().orc_welford_var_estimate(self.n, _p(self.M), C.c_int64(self.mu.size), _p(out))
        return out


class WelfordCov:
    def __init__(self, D):
        self.D = D
        self.n = C.c_int64(0)
        self.mu = np.zeros(D)
        self.M = np.zeros((D, D), order="F")

    def push(self, s):
        s = np.ascontiguousarray(s, dtype=np.float64)
        lib().orc_welford_cov_push(C.byref(self.n), _p(self.mu), _p(self.M), self.D, _p(s))

    def estimate(self):
        out = np.zeros((self.D, self.D), order="F")
        lib().orc_welford_cov_estimate(self.n, _p(self.M), self.D, _p(out))
        return out


def set_partial_refresh(alpha: float):
    """PartialMomentumRefreshment(alpha) for subsequent hmc/nuts transitions (0 = full refresh)."""
    lib().orc_set_partial_refresh(C.c_double(alpha))


def set_tempering(alpha: float):
    """TemperedLeapfrog(eps, alpha) as the integrator of subsequent hmc / multinomial / nuts transitions (0 = plain Leapfrog)."""
    lib().orc_set_tempering(C.c_double(alpha))


def stan_windows(n_adapts, init_buffer=75, term_buffer=50, window_size=25):
    ws, we = C.c_int32(0), C.c_int32(0)
    splits = (C.c_int32 * 64)()
    n = lib().orc_stan_windows(init_buffer, term_buffer, window_size, n_adapts, C.byref(ws), C.byref(we), splits)
    return ws.value, we.value, [splits[i] for i in range(n)]
