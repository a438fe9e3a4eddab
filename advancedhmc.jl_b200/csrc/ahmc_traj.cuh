// ahmc_traj.cuh -- the fused n-step leapfrog trajectory of ONE chain group, shared by K1 (`step`,
// src/integrator.jl:216-265) and K2 (static `transition`, src/trajectory.jl:271-300).
//
// Two code paths (see DESIGN.md "K1"):
//
//  * EXACT (every model x metric): per step the reference's op sequence with FMAs, energies reduced
//    by warp shuffles, and the reference's `isfinite(z)` test (hamiltonian.jl:141-142).  A non-finite
//    chain stops on its own and its phase point AT the break step is what is handed to `done`
//    (integrator.jl:252-258 returns the non-finite z).
//
//  * FAST (separable Gaussian targets STD_NORMAL / DIAG_GAUSS with Unit / Diag metric, no tempering):
//    state in shifted coordinates x = theta - m, the two half kicks of consecutive steps merged,
//    per-coordinate constants a = eps*Minv, b = eps/s^2 precomputed: a step is 2 dependent DFMAs per
//    coordinate and no reduction.  The reference's per-step `isfinite` control flow stays exact by a
//    magnitude argument: with K = (1+max|a|)(1+max|b|) the sup-norm of (x, r) grows by at most K per
//    step, so checking the exponent fields of (x, r) against 2^200 every floor(100/log2 K) steps
//    PROVES that every intermediate phase point, energies included, was finite (|x|,|r| < 2^300,
//    parameters < 2^100, so squares summed over D < 2^31 coordinates stay below 2^731).  A chain that
//    fails a check, or whose parameters are outside the proof's range, is re-run by the exact path in
//    the same launch.
//
// The caller supplies a functor F with
//    void init(double (&th)[E], double (&r)[E], double (&g)[E])   -- (re)materialise the start state
//    void done(th, r, g, dr, lp, lk, fin, steps)                  -- consume the end state (stores)
// `done` is called exactly once per valid chain, by all lanes of the chain's group.
#pragma once
#include "ahmc_device.cuh"

namespace ahmc {

template <int MODEL, int METRIC>
struct FastCapable {
    static constexpr bool value = (MODEL == AHMC_MODEL_STD_NORMAL || MODEL == AHMC_MODEL_DIAG_GAUSS) &&
                                  (METRIC == AHMC_METRIC_UNIT || METRIC == AHMC_METRIC_DIAG);
};

template <int MODEL, int METRIC, int G, int E, class F>
__device__ __forceinline__ void run_trajectory(const ModelDev& model, const MetricDev& metric, int D,
                                               long long chain, bool valid, int l, double* xs, double eps, int n,
                                               double temper_alpha, uint32_t flags, F& f) {
    bool need_exact = valid;

    if constexpr (FastCapable<MODEL, METRIC>::value) {
        const bool fast_on = !(flags & AHMC_FLAG_EXACT_CHECKS) && !(temper_alpha > 0.0);
        if (fast_on) {
            constexpr int T200 = expo_bits(200), T100 = expo_bits(100), T50 = expo_bits(50);
            double x[E], r[E], ca[E], cb[E], mu[E];
            // |eps| must be in [2^-100, 2^50] for the proof below and for the 1/eps rescaling of the last step
            const double inv_eps = 1.0 / eps;
            bool suspicious = big_d(eps, T50) | big_d(inv_eps, T100);
            double Amax = 0.0, Bmax = 0.0;
            const double he = 0.5 * eps;
            {
                double g0[E];
                f.init(x, r, g0);
                const double* pMi = (METRIC == AHMC_METRIC_DIAG) ? metric.Minv + metric.chain_stride * chain + l : nullptr;
                const double* pW = (MODEL == AHMC_MODEL_DIAG_GAUSS) ? model.p1 + l : nullptr;
                const double* pMu = (MODEL == AHMC_MODEL_DIAG_GAUSS) ? model.p0 + l : nullptr;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const bool in = (l + G * e) < D;
                    double mi = (METRIC == AHMC_METRIC_DIAG) ? (in ? __ldg(pMi + G * e) : 0.0) : (in ? 1.0 : 0.0);
                    double wi = (MODEL == AHMC_MODEL_DIAG_GAUSS) ? (in ? __ldg(pW + G * e) : 0.0) : (in ? 1.0 : 0.0);
                    mu[e] = (MODEL == AHMC_MODEL_DIAG_GAUSS) ? (in ? __ldg(pMu + G * e) : 0.0) : 0.0;
                    suspicious |= big_d(mi, T100) | big_d(wi, T100) | big_d(mu[e], T200) | big_d(x[e], T200) |
                                  big_d(r[e], T200) | big_d(g0[e], T200);
                    ca[e] = eps * mi;
                    cb[e] = eps * wi;
                    Amax = fmax(Amax, fabs(ca[e]));
                    Bmax = fmax(Bmax, fabs(cb[e]));
                    x[e] = x[e] - mu[e];           // shifted coordinate
                    r[e] = fma(-he, g0[e], r[e]);  // first half kick uses the CACHED gradient (integrator.jl:237)
                }
            }
            Amax = Grp<G>::max(Amax);
            Bmax = Grp<G>::max(Bmax);
            // K = (1+A)(1+B) bounds the per-step growth of max(|x|,|r|); log2(K) <= exponent(K) + 1
            const double K = (1.0 + Amax) * (1.0 + Bmax);
            int kcheck = n;
            {
                const int ek = ((__double2hiint(K) >> 20) & 0x7ff) - 1023 + 1;  // K >= 1: ek >= 1
                if (ek > 100 || !(K >= 1.0)) suspicious = true;                // also catches NaN / Inf
                const int kk = 100 / (ek < 1 ? 1 : ek);
                kcheck = kk < 1 ? 1 : kk;
            }
            // n-1 x (drift + merged full kick), magnitude check every kcheck steps
            int remaining = n - 1;
            while (remaining > 0) {
                const int c = remaining < kcheck ? remaining : kcheck;
                for (int j = 0; j < c; ++j) {
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        x[e] = fma(ca[e], r[e], x[e]);
                        r[e] = fma(-cb[e], x[e], r[e]);
                    }
                }
                remaining -= c;
#pragma unroll
                for (int e = 0; e < E; ++e) suspicious |= big_d(x[e], T200) | big_d(r[e], T200);
            }
            // last step: drift, gradient, half kick, energies.  g = x*w and dH/dr = Minv*r are recovered from the
            // per-coordinate constants as (x*b)/eps and (r*a)/eps (one extra rounding, ~1e-16 relative)
            double g[E], dr[E];
            double lp_part = 0.0, lk_part = 0.0;
#pragma unroll
            for (int e = 0; e < E; ++e) {
                x[e] = fma(ca[e], r[e], x[e]);
                g[e] = (MODEL == AHMC_MODEL_DIAG_GAUSS) ? (x[e] * cb[e]) * inv_eps : ((l + G * e) < D ? x[e] : 0.0);
                r[e] = fma(-he, g[e], r[e]);
                suspicious |= big_d(x[e], T200) | big_d(r[e], T200);
                lp_part = fma(x[e], g[e], lp_part);
                dr[e] = (METRIC == AHMC_METRIC_DIAG) ? (r[e] * ca[e]) * inv_eps : r[e];
                lk_part = fma(r[e], dr[e], lk_part);
                x[e] = x[e] + mu[e];  // back to theta
            }
            suspicious = Grp<G>::any(suspicious);
            const double lp = fma(-0.5, Grp<G>::sum(lp_part), model.c0);
            const double lk = -0.5 * Grp<G>::sum(lk_part);
            need_exact = valid && suspicious;
            if (valid && !suspicious) f.done(x, r, g, dr, lp, lk, true, n);  // finite by the magnitude proof
        }
    }

    if (!__any_sync(FULL, need_exact)) return;

    // ------------------------------------------------------------------ EXACT path
    ModelOps<MODEL, G, E> mo;
    MetricOps<METRIC, G, E> me;
    mo.load(model, l, D);
    me.load(metric, chain, l, D);
    ChainState<E> s;
    f.init(s.th, s.r, s.g);
    s.lp = 0.0;
    s.lk = 0.0;
    double dr[E];
    const double sa = temper_alpha > 0.0 ? sqrt(temper_alpha) : 1.0;
    bool active = need_exact;
    for (int i = 1; i <= n; ++i) {
        double t1 = 1.0, t2 = 1.0;
        if (temper_alpha > 0.0) {  // integrator.jl:198-209
            t1 = (2 * (i - 1) + 1 <= n) ? sa : 1.0 / sa;
            t2 = (2 * (i - 1) + 2 <= n) ? sa : 1.0 / sa;
        }
        const bool fin = leapfrog_step<MODEL, METRIC, G, E>(s, mo, me, eps, dr, xs, l, t1, t2);
        if (active && (!fin || i == n)) {
            f.done(s.th, s.r, s.g, dr, s.lp, s.lk, fin, i);
            active = false;
        }
        if (!__any_sync(FULL, active)) break;
    }
}

}  // namespace ahmc
