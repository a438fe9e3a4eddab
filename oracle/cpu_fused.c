/*
 * cpu_fused.c -- the "good CPU implementation" timing baseline: same arithmetic as orc_leapfrog
 * (integrator.jl:216-265 per chain) but fused per chain, FMA allowed, chains split over OpenMP threads.
 *
 * *** TEST / BENCH INFRASTRUCTURE ONLY (lives under oracle/): never on the product path. ***
 * Used by bench.py's cpu_baseline / --impl reference legs.  Results are checked against
 * orc_leapfrog in tests/test_oracle.py.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "ahmc_oracle.h"

static void fused_separable_chain(const orc_model* m, const orc_metric* me, int D, int64_t c, double eps, int n_steps,
                                  const double* th0, const double* r0, const double* g0, double* th, double* r,
                                  double* g, double* lp_out, double* lk_out, const double* wv) {
    const double* mu = m->kind == ORC_MODEL_DIAG_GAUSS ? m->p0 : NULL;
    const double* Mi = me->kind == ORC_METRIC_DIAG ? me->Minv + me->chain_stride * c : NULL;
    double he = eps / 2;
    double lp = 0, lk = 0;
    for (int d = 0; d < D; ++d) {
        th[d] = th0[d];
        r[d] = r0[d];
        g[d] = g0[d];
    }
    for (int i = 0; i < n_steps; ++i) {
        lp = 0;
        lk = 0;
        for (int d = 0; d < D; ++d) {
            double rr = r[d] - he * g[d];
            double mi = Mi ? Mi[d] : 1.0;
            double t = th[d] + eps * (mi * rr);
            double diff = mu ? (mu[d] - t) : -t;
            double w = wv[d];
            double gg = -(diff * w);
            lp += -(diff * diff * w) / 2;
            rr = rr - he * gg;
            lk += rr * rr * mi;
            th[d] = t;
            r[d] = rr;
            g[d] = gg;
        }
        lp = lp + m->c0;
        lk = -lk / 2;
        if (!isfinite(lp) || !isfinite(lk)) break; /* element non-finiteness implies these for these models */
    }
    *lp_out = isfinite(lp) ? lp : -INFINITY;
    *lk_out = isfinite(lk) ? lk : -INFINITY;
}

void orc_leapfrog_omp(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                      const double* eps_chain, int32_t n_steps, const orc_phasepoint* z_in,
                      const orc_phasepoint* z_out, int32_t n_threads) {
    int separable = (m->kind == ORC_MODEL_STD_NORMAL || m->kind == ORC_MODEL_DIAG_GAUSS) &&
                    (me->kind == ORC_METRIC_UNIT || me->kind == ORC_METRIC_DIAG) && n_steps > 0;
#ifdef _OPENMP
    if (n_threads > 0) omp_set_num_threads(n_threads);
#endif
    if (separable) {
        double* wv = (double*)malloc(sizeof(double) * (size_t)D); /* 1/s^2 hoisted out of the step loop */
        for (int d = 0; d < D; ++d) wv[d] = m->kind == ORC_MODEL_DIAG_GAUSS ? 1.0 / (m->p1[d] * m->p1[d]) : 1.0;
#pragma omp parallel for schedule(static)
        for (int64_t c = 0; c < N; ++c) {
            double e = eps_chain ? eps_chain[c] : eps;
            fused_separable_chain(m, me, D, c, e, n_steps, z_in->theta + z_in->ld * c, z_in->r + z_in->ld * c,
                                  z_in->lp_gradient + z_in->ld * c, z_out->theta + z_out->ld * c,
                                  z_out->r + z_out->ld * c, z_out->lp_gradient + z_out->ld * c, &z_out->lp_value[c],
                                  &z_out->lk_value[c], wv);
        }
        free(wv);
        return;
    }
    /* generic: chunk chains over threads, each chunk through the scalar oracle */
    int nt = 1;
#ifdef _OPENMP
    nt = omp_get_max_threads();
#endif
    int64_t chunk = (N + nt - 1) / nt;
#pragma omp parallel for schedule(static)
    for (int t = 0; t < nt; ++t) {
        int64_t c0 = t * chunk, c1 = c0 + chunk > N ? N : c0 + chunk;
        if (c0 >= c1) continue;
        orc_phasepoint a = *z_in, b = *z_out;
        a.theta += a.ld * c0; a.r += a.ld * c0; a.lp_gradient += a.ld * c0; a.lp_value += c0; a.lk_value += c0;
        if (a.lk_gradient) a.lk_gradient += a.ld * c0;
        b.theta += b.ld * c0; b.r += b.ld * c0; b.lp_gradient += b.ld * c0; b.lp_value += c0; b.lk_value += c0;
        if (b.lk_gradient) b.lk_gradient += b.ld * c0;
        orc_metric me2 = *me;
        if (me2.kind == ORC_METRIC_DIAG && me2.chain_stride) me2.Minv += me2.chain_stride * c0;
        orc_leapfrog(m, &me2, D, c1 - c0, eps, eps_chain ? eps_chain + c0 : NULL, n_steps, 0.0, &a, &b, NULL, NULL, 0);
    }
}
