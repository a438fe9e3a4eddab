/*
 * ahmc_oracle.c -- CPU restatement of AdvancedHMC.jl's vectorised leapfrog / HMC / NUTS path.
 *
 * *** TEST INFRASTRUCTURE ONLY -- see ahmc_oracle.h for the rules and the parity-pin status
 * ("parity unpinned" against the Julia reference itself: it cannot run in this image). ***
 *
 * Written op-for-op after the reference (citations: /root/reference/<file>:<line>), scalar loops,
 * no FMA contraction (build with -ffp-contract=off; Julia does not fuse `a - b .* c`).
 */
#include "ahmc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NEG_INF (-INFINITY)

/* Julia's min/max propagate NaN (Base.min); C fmin does not. */
static double jl_min(double a, double b) {
    if (isnan(a) || isnan(b)) return NAN;
    return a < b ? a : b;
}
static double jl_max(double a, double b) {
    if (isnan(a) || isnan(b)) return NAN;
    return a > b ? a : b;
}

/* ------------------------------------------------------------------------------------------ */
/* user closure  (hamiltonian.jl:45-48 calls h.dlp/dth(theta) -> (lp, grad))                    */
/* ------------------------------------------------------------------------------------------ */
void orc_logp_grad(const orc_model* m, const double* th, double* lp_out, double* grad) {
    const int D = m->D;
    double lp = 0.0;
    switch (m->kind) {
        case ORC_MODEL_STD_NORMAL: {
            for (int d = 0; d < D; ++d) {
                lp += -(th[d] * th[d]) / 2; /* -abs2(x)/2, summed along dim 1 */
                grad[d] = -th[d];
            }
            break;
        }
        case ORC_MODEL_DIAG_GAUSS: { /* test/common.jl:40-56, with the true gradient (m-x)/s^2 */
            const double* mu = m->p0;
            const double* s = m->p1;
            for (int d = 0; d < D; ++d) {
                double g = mu[d] - th[d];
                double s2 = s[d] * s[d];
                lp += -((g * g) / s2) / 2;
                grad[d] = g / s2;
            }
            break;
        }
        case ORC_MODEL_DENSE_GAUSS: {
            const double* mu = m->p0;
            const double* P = m->p1; /* column-major D x D */
            double q = 0.0;
            /* grad = -(P * diff): column-sweep gemv like BLAS */
            for (int d = 0; d < D; ++d) grad[d] = 0.0;
            for (int k = 0; k < D; ++k) {
                double dk = th[k] - mu[k];
                for (int d = 0; d < D; ++d) grad[d] += P[d + (int64_t)D * k] * dk;
            }
            for (int d = 0; d < D; ++d) {
                q += (th[d] - mu[d]) * grad[d];
                grad[d] = -grad[d];
            }
            lp = -q / 2;
            break;
        }
        case ORC_MODEL_FUNNEL: { /* th1~N(0,3), th_i~N(0,exp(th1/2)) */
            double v = th[0];
            double ev = exp(-v);
            double S = 0.0;
            for (int d = 1; d < D; ++d) {
                S += th[d] * th[d] * ev;
                grad[d] = -th[d] * ev;
            }
            lp = -(v * v) / 18 - (S + (D - 1) * v) / 2;
            grad[0] = -v / 9 + (S - (D - 1)) / 2;
            break;
        }
        default:
            lp = NAN;
    }
    *lp_out = lp + m->c0;
}

/* ------------------------------------------------------------------------------------------ */
/* metric ops                                                                                  */
/* ------------------------------------------------------------------------------------------ */
void orc_dHdr(const orc_metric* me, int32_t D, int64_t c, const double* r, double* out) {
    switch (me->kind) {
        case ORC_METRIC_UNIT: /* hamiltonian.jl:50  copy(r) */
            for (int d = 0; d < D; ++d) out[d] = r[d];
            break;
        case ORC_METRIC_DIAG: { /* hamiltonian.jl:51-59  Minv .* r */
            const double* Mi = me->Minv + me->chain_stride * c;
            for (int d = 0; d < D; ++d) out[d] = Mi[d] * r[d];
            break;
        }
        case ORC_METRIC_DENSE: { /* hamiltonian.jl:60-68  Minv * r */
            const double* Mi = me->Minv;
            for (int d = 0; d < D; ++d) out[d] = 0.0;
            for (int k = 0; k < D; ++k) {
                double rk = r[k];
                for (int d = 0; d < D; ++d) out[d] += Mi[d + (int64_t)D * k] * rk;
            }
            break;
        }
    }
}

double orc_neg_kinetic(const orc_metric* me, int32_t D, int64_t c, const double* r) {
    double s = 0.0;
    switch (me->kind) {
        case ORC_METRIC_UNIT: /* hamiltonian.jl:155-165  -sum(abs2, r)/2 */
            for (int d = 0; d < D; ++d) s += r[d] * r[d];
            return -s / 2;
        case ORC_METRIC_DIAG: { /* hamiltonian.jl:167-177  -sum(abs2.(r) .* Minv)/2 */
            const double* Mi = me->Minv + me->chain_stride * c;
            for (int d = 0; d < D; ++d) s += (r[d] * r[d]) * Mi[d];
            return -s / 2;
        }
        case ORC_METRIC_DENSE: { /* hamiltonian.jl:179-184  mul!(_temp, Minv, r); -dot(r,_temp)/2 */
            double* tmp = (double*)malloc(sizeof(double) * (size_t)D);
            orc_dHdr(me, D, c, r, tmp);
            for (int d = 0; d < D; ++d) s += r[d] * tmp[d];
            free(tmp);
            return -s / 2;
        }
    }
    return NAN;
}

void orc_rand_momentum(const orc_metric* me, int32_t D, int64_t c, const double* z, double* r) {
    switch (me->kind) {
        case ORC_METRIC_UNIT: /* metric.jl:290-298 */
            for (int d = 0; d < D; ++d) r[d] = z[d];
            break;
        case ORC_METRIC_DIAG: { /* metric.jl:300-309  r ./= sqrtMinv, sqrtMinv = sqrt.(Minv) (:61-63) */
            const double* Mi = me->Minv + me->chain_stride * c;
            for (int d = 0; d < D; ++d) r[d] = z[d] / sqrt(Mi[d]);
            break;
        }
        case ORC_METRIC_DENSE: { /* metric.jl:311-320  ldiv!(cholMinv, r): back substitution with U */
            const double* U = me->cholU;
            for (int d = 0; d < D; ++d) r[d] = z[d];
            for (int i = D - 1; i >= 0; --i) {
                double x = r[i];
                for (int k = i + 1; k < D; ++k) x -= U[i + (int64_t)D * k] * r[k];
                r[i] = x / U[i + (int64_t)D * i];
            }
            break;
        }
    }
}

/* PhasePoint ctor maps non-finite VALUES to -Inf (hamiltonian.jl:95-104) */
static double map_nonfinite(double v) { return isfinite(v) ? v : NEG_INF; }

static int all_finite(const double* x, int D) {
    for (int d = 0; d < D; ++d)
        if (!isfinite(x[d])) return 0;
    return 1;
}

void orc_make_phasepoint(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, const orc_phasepoint* z) {
    double* grad = (double*)malloc(sizeof(double) * (size_t)D);
    for (int64_t c = 0; c < N; ++c) {
        const double* th = z->theta + z->ld * c;
        const double* r = z->r + z->ld * c;
        double lp;
        orc_logp_grad(m, th, &lp, grad);
        for (int d = 0; d < D; ++d) z->lp_gradient[z->ld * c + d] = -grad[d]; /* hamiltonian.jl:47 */
        z->lp_value[c] = map_nonfinite(lp);
        z->lk_value[c] = map_nonfinite(orc_neg_kinetic(me, D, c, r));
        if (z->lk_gradient) orc_dHdr(me, D, c, r, z->lk_gradient + z->ld * c);
    }
    free(grad);
}

/* ------------------------------------------------------------------------------------------ */
/* leapfrog  (integrator.jl:216-265)                                                           */
/* ------------------------------------------------------------------------------------------ */
/* one leapfrog step of one chain, in place.  g holds -grad log pi (DualValue.gradient).
 * returns isfinite(z) per hamiltonian.jl:141-142 */
static int lf_one_step(const orc_model* m, const orc_metric* me, int D, int64_t c, double eps, int i,
                       int n_steps, double temper_alpha, double* th, double* r, double* g, double* lp,
                       double* lk, double* dr /* scratch D, receives dHdr(r_final) */, double* grad /* scratch */) {
    /* temper (integrator.jl:198-209): first half steps multiply by sqrt(alpha), later divide */
    if (temper_alpha > 0) {
        int i_temper = 2 * (i - 1) + 1;
        double sa = sqrt(temper_alpha);
        for (int d = 0; d < D; ++d) r[d] = (i_temper <= n_steps) ? r[d] * sa : r[d] / sa;
    }
    /* r = r - eps/2 .* gradient   (integrator.jl:237) */
    double he = eps / 2;
    for (int d = 0; d < D; ++d) r[d] = r[d] - he * g[d];
    /* dr = dHdr(h, r); th = th + eps .* dr   (integrator.jl:239-240) */
    orc_dHdr(me, D, c, r, dr);
    for (int d = 0; d < D; ++d) th[d] = th[d] + eps * dr[d];
    /* (value, gradient) = dHdth(h, th)   (integrator.jl:242; hamiltonian.jl:45-48) */
    double v;
    orc_logp_grad(m, th, &v, grad);
    for (int d = 0; d < D; ++d) g[d] = -grad[d];
    /* r = r - eps/2 .* gradient   (integrator.jl:243) */
    for (int d = 0; d < D; ++d) r[d] = r[d] - he * g[d];
    if (temper_alpha > 0) {
        int i_temper = 2 * (i - 1) + 2;
        double sa = sqrt(temper_alpha);
        for (int d = 0; d < D; ++d) r[d] = (i_temper <= n_steps) ? r[d] * sa : r[d] / sa;
    }
    /* z = phasepoint(h, th, r; lp=DualValue(value, gradient))  (integrator.jl:247; hamiltonian.jl:115-119) */
    double k = orc_neg_kinetic(me, D, c, r);
    orc_dHdr(me, D, c, r, dr);
    *lp = map_nonfinite(v);
    *lk = map_nonfinite(k);
    /* isfinite(z) (hamiltonian.jl:141-142) */
    return isfinite(*lp) && all_finite(g, D) && isfinite(*lk) && all_finite(dr, D);
}

static void copy_pp(int D, int64_t N, const orc_phasepoint* a, const orc_phasepoint* b) {
    if (a->theta != b->theta)
        for (int64_t c = 0; c < N; ++c) memcpy(b->theta + b->ld * c, a->theta + a->ld * c, sizeof(double) * (size_t)D);
    if (a->r != b->r)
        for (int64_t c = 0; c < N; ++c) memcpy(b->r + b->ld * c, a->r + a->ld * c, sizeof(double) * (size_t)D);
    if (a->lp_gradient != b->lp_gradient)
        for (int64_t c = 0; c < N; ++c)
            memcpy(b->lp_gradient + b->ld * c, a->lp_gradient + a->ld * c, sizeof(double) * (size_t)D);
    if (a->lp_value != b->lp_value) memcpy(b->lp_value, a->lp_value, sizeof(double) * (size_t)N);
    if (a->lk_value != b->lk_value) memcpy(b->lk_value, a->lk_value, sizeof(double) * (size_t)N);
    if (a->lk_gradient && b->lk_gradient && a->lk_gradient != b->lk_gradient)
        for (int64_t c = 0; c < N; ++c)
            memcpy(b->lk_gradient + b->ld * c, a->lk_gradient + a->ld * c, sizeof(double) * (size_t)D);
}

static void leapfrog_impl(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                          const double* eps_chain, int32_t n_steps_signed, double temper_alpha,
                          const orc_phasepoint* z_in, const orc_phasepoint* z_out, const orc_phasepoint* traj,
                          int64_t step_stride, uint32_t* status, int32_t* steps_done, int compat_break_all) {
    int n_steps = n_steps_signed < 0 ? -n_steps_signed : n_steps_signed; /* integrator.jl:224 */
    int fwd = n_steps_signed > 0;                                        /* integrator.jl:221 */
    double* th = (double*)malloc(sizeof(double) * (size_t)D * 5);
    double *r = th + D, *g = th + 2 * D, *dr = th + 3 * D, *grad = th + 4 * D;
    /* working state lives in z_out (or in a private copy when only a trajectory is wanted) */
    orc_phasepoint w;
    double* own = NULL;
    if (z_out) {
        w = *z_out;
        copy_pp(D, N, z_in, z_out);
    } else {
        own = (double*)malloc(sizeof(double) * ((size_t)D * N * 3 + (size_t)N * 2));
        w.theta = own;
        w.r = own + (size_t)D * N;
        w.lp_gradient = own + (size_t)D * N * 2;
        w.lp_value = own + (size_t)D * N * 3;
        w.lk_value = w.lp_value + N;
        w.lk_gradient = NULL;
        w.ld = D;
        copy_pp(D, N, z_in, &w);
    }
    uint8_t* active = (uint8_t*)malloc((size_t)N);
    for (int64_t c = 0; c < N; ++c) {
        active[c] = 1;
        if (status) status[c] = 0;
        if (steps_done) steps_done[c] = 0;
    }
    for (int i = 1; i <= n_steps; ++i) {
        int any_nonfinite = 0;
        for (int64_t c = 0; c < N; ++c) {
            if (!active[c]) continue;
            double e = eps_chain ? eps_chain[c] : eps;
            e = fwd ? e : -e; /* integrator.jl:226 */
            memcpy(th, w.theta + w.ld * c, sizeof(double) * (size_t)D);
            memcpy(r, w.r + w.ld * c, sizeof(double) * (size_t)D);
            memcpy(g, w.lp_gradient + w.ld * c, sizeof(double) * (size_t)D);
            double lp, lk;
            int fin = lf_one_step(m, me, D, c, e, i, n_steps, temper_alpha, th, r, g, &lp, &lk, dr, grad);
            memcpy(w.theta + w.ld * c, th, sizeof(double) * (size_t)D);
            memcpy(w.r + w.ld * c, r, sizeof(double) * (size_t)D);
            memcpy(w.lp_gradient + w.ld * c, g, sizeof(double) * (size_t)D);
            w.lp_value[c] = lp;
            w.lk_value[c] = lk;
            if (w.lk_gradient) memcpy(w.lk_gradient + w.ld * c, dr, sizeof(double) * (size_t)D);
            if (steps_done) steps_done[c] = i;
            if (traj) { /* res[i] = z  (integrator.jl:249-251) */
                int64_t o = (int64_t)(i - 1) * step_stride + traj->ld * c;
                memcpy(traj->theta + o, th, sizeof(double) * (size_t)D);
                memcpy(traj->r + o, r, sizeof(double) * (size_t)D);
                memcpy(traj->lp_gradient + o, g, sizeof(double) * (size_t)D);
                if (traj->lk_gradient) memcpy(traj->lk_gradient + o, dr, sizeof(double) * (size_t)D);
                traj->lp_value[(int64_t)(i - 1) * N + c] = lp;
                traj->lk_value[(int64_t)(i - 1) * N + c] = lk;
            }
            if (!fin) { /* integrator.jl:252-258 */
                any_nonfinite = 1;
                if (status) status[c] |= 1u;
                if (!compat_break_all) active[c] = 0;
            }
        }
        if (compat_break_all && any_nonfinite) break;
    }
    free(active);
    free(th);
    free(own);
}

void orc_leapfrog(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                  const double* eps_chain, int32_t n_steps, double temper_alpha, const orc_phasepoint* z_in,
                  const orc_phasepoint* z_out, uint32_t* status, int32_t* steps_done, int compat_break_all) {
    leapfrog_impl(m, me, D, N, eps, eps_chain, n_steps, temper_alpha, z_in, z_out, NULL, 0, status, steps_done,
                  compat_break_all);
}

void orc_leapfrog_trajectory(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                             const double* eps_chain, int32_t n_steps, double temper_alpha,
                             const orc_phasepoint* z_in, const orc_phasepoint* traj, int64_t step_stride,
                             int32_t* steps_done, int compat_break_all) {
    leapfrog_impl(m, me, D, N, eps, eps_chain, n_steps, temper_alpha, z_in, NULL, traj, step_stride, NULL,
                  steps_done, compat_break_all);
}

static double g_partial_alpha = 0.0;
void orc_set_partial_refresh(double alpha) { g_partial_alpha = alpha; }
/* TemperedLeapfrog(eps, alpha) as the integrator of the transitions below (integrator.jl:174-209): every `step` call a
   transition makes -- the static trajectory (trajectory.jl:337), each leg of the multinomial one (:374-376), each NUTS
   leaf (:640, n_steps = +-1) -- tempers by its own n_steps. */
static double g_temper_alpha = 0.0;
void orc_set_tempering(double alpha) { g_temper_alpha = alpha > 0 ? alpha : 0.0; }
/* refresh(rng, ref, h, z) (hamiltonian.jl:213-220 full, :243-254 partial) for chain c */
static void refresh_momentum(const orc_metric* me, int D, int64_t c, const double* z_tape, const double* r_prev, double* r) {
    orc_rand_momentum(me, D, c, z_tape, r);
    if (g_partial_alpha != 0.0) {
        double a = g_partial_alpha, b = sqrt(1 - a * a);
        for (int d = 0; d < D; ++d) r[d] = a * r_prev[d] + b * r[d];
    }
}

/* ------------------------------------------------------------------------------------------ */
/* static HMC transition (sampler.jl:48-58; trajectory.jl:271-300, :312-340, :863-880)          */
/* ------------------------------------------------------------------------------------------ */
void orc_hmc_transition(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                        const double* eps_chain, int32_t n_steps, const double* normal_tape,
                        const double* exp_tape, const orc_phasepoint* z_in, const orc_phasepoint* z_out,
                        const orc_stats* st, int compat_break_all) {
    size_t DN = (size_t)D * N;
    /* z0 = refresh(rng, FullMomentumRefreshment, h, z) (hamiltonian.jl:213-220): new r, lk recomputed,
       lp recomputed from theta (same value; quirk Q2) */
    double* buf = (double*)malloc(sizeof(double) * (DN * 6 + (size_t)N * 4));
    orc_phasepoint z0 = {buf, buf + DN, buf + DN * 6, buf + DN * 2, buf + DN * 6 + N, NULL, D};
    orc_phasepoint z1 = {buf + DN * 3, buf + DN * 4, buf + DN * 6 + 2 * N, buf + DN * 5, buf + DN * 6 + 3 * N, NULL, D};
    for (int64_t c = 0; c < N; ++c) {
        memcpy(z0.theta + D * c, z_in->theta + z_in->ld * c, sizeof(double) * (size_t)D);
        if (normal_tape)
            refresh_momentum(me, D, c, normal_tape + (size_t)D * c, z_in->r + z_in->ld * c, z0.r + D * c);
        else
            memcpy(z0.r + D * c, z_in->r + z_in->ld * c, sizeof(double) * (size_t)D);
    }
    orc_make_phasepoint(m, me, D, N, &z0);
    /* z' = step(integrator, h, z, nsteps)  (trajectory.jl:337) */
    leapfrog_impl(m, me, D, N, eps, eps_chain, n_steps, g_temper_alpha, &z0, &z1, NULL, 0, NULL, NULL, compat_break_all);
    for (int64_t c = 0; c < N; ++c) {
        double H0 = -(z0.lp_value[c] + z0.lk_value[c]); /* energy(z) hamiltonian.jl:149,194 */
        double H1 = -(z1.lp_value[c] + z1.lk_value[c]);
        /* mh_accept_ratio (trajectory.jl:863-880) */
        int accept = H1 < H0 + exp_tape[c];
        double alpha = jl_min(1.0, exp(H0 - H1));
        /* accept_phasepoint! (trajectory.jl:312-332) then flip (trajectory.jl:283) */
        const orc_phasepoint* src = accept ? &z1 : &z0;
        for (int d = 0; d < D; ++d) {
            z_out->theta[z_out->ld * c + d] = src->theta[D * c + d];
            z_out->r[z_out->ld * c + d] = -src->r[D * c + d];
            z_out->lp_gradient[z_out->ld * c + d] = src->lp_gradient[D * c + d];
        }
        z_out->lp_value[c] = src->lp_value[c];
        z_out->lk_value[c] = src->lk_value[c];
        if (z_out->lk_gradient) /* lk.gradient is carried over un-negated (trajectory.jl:283 passes z.lk) */
            orc_dHdr(me, D, c, src->r + D * c, z_out->lk_gradient + z_out->ld * c);
        double H = -(src->lp_value[c] + src->lk_value[c]);
        if (st) {
            if (st->n_steps) st->n_steps[c] = n_steps;
            if (st->is_accept) st->is_accept[c] = (uint8_t)accept;
            if (st->acceptance_rate) st->acceptance_rate[c] = alpha;
            if (st->log_density) st->log_density[c] = src->lp_value[c];
            if (st->hamiltonian_energy) st->hamiltonian_energy[c] = H;
            if (st->hamiltonian_energy_error) st->hamiltonian_energy_error[c] = H - H0;
            if (st->numerical_error) st->numerical_error[c] = (uint8_t)!isfinite(H1); /* per chain; ref ORs over chains */
        }
    }
    free(buf);
}

/* ------------------------------------------------------------------------------------------ */
/* static HMC transition, MultinomialTS (trajectory.jl:344-390)                                 */
/* ------------------------------------------------------------------------------------------ */
void orc_hmc_multinomial_transition(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                                    const double* eps_chain, int32_t n_steps, int32_t n_steps_fwd,
                                    const double* normal_tape, const double* unif_tape, const orc_phasepoint* z_in,
                                    const orc_phasepoint* z_out, const orc_stats* st) {
    const int nf_req = n_steps_fwd, nb_req = n_steps - n_steps_fwd;
    const int L = n_steps + 1;
    /* one chain at a time (per-chain break = the reference applied to a single chain) */
    double* buf = (double*)malloc(sizeof(double) * ((size_t)D * 3 * (size_t)(L + 1) + (size_t)L * 4 + 8));
    double* TH = buf;                       /* (L) x D : trajectory order */
    double* R = TH + (size_t)D * L;
    double* G = R + (size_t)D * L;
    double* LP = G + (size_t)D * L;
    double* LK = LP + L;
    double* W = LK + L;
    double* scratch = W + L; /* unused tail */
    (void)scratch;
    double* th = (double*)malloc(sizeof(double) * (size_t)D * 5);
    double *r = th + D, *g = th + 2 * D, *dr = th + 3 * D, *grad = th + 4 * D;
    for (int64_t c = 0; c < N; ++c) {
        double e = eps_chain ? eps_chain[c] : eps;
        /* z = refresh(...) */
        double th0[1024], r0[1024], g0[1024];
        if (D > 1024) abort();
        memcpy(th0, z_in->theta + z_in->ld * c, sizeof(double) * (size_t)D);
        if (normal_tape) refresh_momentum(me, D, c, normal_tape + (size_t)D * c, z_in->r + z_in->ld * c, r0);
        else memcpy(r0, z_in->r + z_in->ld * c, sizeof(double) * (size_t)D);
        double lp0;
        orc_logp_grad(m, th0, &lp0, grad);
        for (int d = 0; d < D; ++d) g0[d] = -grad[d];
        lp0 = map_nonfinite(lp0);
        double lk0 = map_nonfinite(orc_neg_kinetic(me, D, c, r0));
        /* backward sweep first so that the arrays end up in trajectory order: reverse(bwd)..., z, fwd... */
        int nb = 0, nf = 0;
        double tmpTH[1], *bTH = (double*)malloc(sizeof(double) * (size_t)D * 3 * (size_t)(nb_req + 1) + 16);
        (void)tmpTH;
        double *bR = bTH + (size_t)D * (nb_req + 1), *bG = bR + (size_t)D * (nb_req + 1);
        double* bLP = (double*)malloc(sizeof(double) * 2 * (size_t)(nb_req + 1));
        double* bLK = bLP + (nb_req + 1);
        memcpy(th, th0, sizeof(double) * (size_t)D); memcpy(r, r0, sizeof(double) * (size_t)D); memcpy(g, g0, sizeof(double) * (size_t)D);
        for (int i = 1; i <= nb_req; ++i) {
            double lp, lk;
            int fin = lf_one_step(m, me, D, c, -e, i, nb_req, g_temper_alpha, th, r, g, &lp, &lk, dr, grad);
            memcpy(bTH + (size_t)D * nb, th, sizeof(double) * (size_t)D);
            memcpy(bR + (size_t)D * nb, r, sizeof(double) * (size_t)D);
            memcpy(bG + (size_t)D * nb, g, sizeof(double) * (size_t)D);
            bLP[nb] = lp; bLK[nb] = lk;
            nb++;
            if (!fin) break;
        }
        int len = 0;
        for (int i = nb - 1; i >= 0; --i, ++len) {
            memcpy(TH + (size_t)D * len, bTH + (size_t)D * i, sizeof(double) * (size_t)D);
            memcpy(R + (size_t)D * len, bR + (size_t)D * i, sizeof(double) * (size_t)D);
            memcpy(G + (size_t)D * len, bG + (size_t)D * i, sizeof(double) * (size_t)D);
            LP[len] = bLP[i]; LK[len] = bLK[i];
        }
        free(bTH); free(bLP);
        memcpy(TH + (size_t)D * len, th0, sizeof(double) * (size_t)D);
        memcpy(R + (size_t)D * len, r0, sizeof(double) * (size_t)D);
        memcpy(G + (size_t)D * len, g0, sizeof(double) * (size_t)D);
        LP[len] = lp0; LK[len] = lk0; len++;
        memcpy(th, th0, sizeof(double) * (size_t)D); memcpy(r, r0, sizeof(double) * (size_t)D); memcpy(g, g0, sizeof(double) * (size_t)D);
        for (int i = 1; i <= nf_req; ++i) {
            double lp, lk;
            int fin = lf_one_step(m, me, D, c, e, i, nf_req, g_temper_alpha, th, r, g, &lp, &lk, dr, grad);
            memcpy(TH + (size_t)D * len, th, sizeof(double) * (size_t)D);
            memcpy(R + (size_t)D * len, r, sizeof(double) * (size_t)D);
            memcpy(G + (size_t)D * len, g, sizeof(double) * (size_t)D);
            LP[len] = lp; LK[len] = lk; len++; nf++;
            if (!fin) break;
        }
        /* weights = -energy.(zs); P = exp.(w .- logsumexp(w)); idx = count(cumsum(P) .< u) + 1 */
        double H0 = -(lp0 + lk0);
        double mx = NEG_INF;
        for (int i = 0; i < len; ++i) { W[i] = LP[i] + LK[i]; if (W[i] > mx) mx = W[i]; }
        double ssum = 0.0;
        for (int i = 0; i < len; ++i) ssum += exp(W[i] - mx);
        double lse = mx + log(ssum);
        double u = unif_tape[c], C = 0.0, asum = 0.0;
        int cnt = 0;
        for (int i = 0; i < len; ++i) {
            C += exp(W[i] - lse);
            if (C < u) cnt++;
            double dH = -W[i] - H0;
            asum += exp(jl_min(0.0, -dH));
        }
        int idx = cnt; /* 0-based index of the (cnt+1)-th point */
        if (idx > len - 1) idx = len - 1;
        double alpha = asum / len;
        for (int d = 0; d < D; ++d) {
            z_out->theta[z_out->ld * c + d] = TH[(size_t)D * idx + d];
            z_out->r[z_out->ld * c + d] = -R[(size_t)D * idx + d]; /* flip (trajectory.jl:283) */
            z_out->lp_gradient[z_out->ld * c + d] = G[(size_t)D * idx + d];
        }
        z_out->lp_value[c] = LP[idx];
        z_out->lk_value[c] = LK[idx];
        if (z_out->lk_gradient) orc_dHdr(me, D, c, R + (size_t)D * idx, z_out->lk_gradient + z_out->ld * c);
        double H = -(LP[idx] + LK[idx]);
        if (st) {
            if (st->n_steps) st->n_steps[c] = n_steps;
            if (st->is_accept) st->is_accept[c] = 1;
            if (st->acceptance_rate) st->acceptance_rate[c] = alpha;
            if (st->log_density) st->log_density[c] = LP[idx];
            if (st->hamiltonian_energy) st->hamiltonian_energy[c] = H;
            if (st->hamiltonian_energy_error) st->hamiltonian_energy_error[c] = H - H0;
            if (st->numerical_error) st->numerical_error[c] = (uint8_t)!isfinite(H);
            if (st->tree_depth) st->tree_depth[c] = idx - nb; /* signed offset of the chosen point from z (test aid) */
        }
        (void)nf;
    }
    free(th);
    free(buf);
}

/* ------------------------------------------------------------------------------------------ */
/* NUTS  (trajectory.jl:626-742), MultinomialTS + GeneralisedNoUTurn, recursive like the ref    */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    double* theta;
    double* r;
    double* g; /* lp.gradient = -grad log pi */
    double lp, lk;
} pp_t;

typedef struct {
    const pp_t* zleft;
    const pp_t* zright;
    double* rho;
    double sum_alpha;
    int n_alpha;
    double dH_max;
} tree_t;

typedef struct {
    const pp_t* zcand;
    double lw; /* MultinomialTS: log weight; SliceTS: number of acceptable candidates n (as a double) */
} sampler_t;

typedef struct {
    int dynamic, numerical;
} term_t;

typedef struct {
    const orc_model* m;
    const orc_metric* me;
    int D;
    int64_t c;
    double eps;
    double delta_max;
    const uint8_t* dirs;
    const double* exps;
    int n_exp;
    /* bump arena */
    double* arena;
    size_t arena_used, arena_cap;
    pp_t* pps;
    size_t pp_used, pp_cap;
    double* scratch; /* 2D */
    int sampler;   /* 0 MultinomialTS, 1 SliceTS (trajectory.jl:102-136) */
    int criterion; /* 0 GeneralisedNoUTurn, 1 ClassicNoUTurn, 2 StrictGeneralisedNoUTurn (trajectory.jl:414-452) */
    double lu;     /* SliceTS slice variable */
} nuts_ctx;

static double* arena_alloc(nuts_ctx* x, size_t n) {
    if (x->arena_used + n > x->arena_cap) abort();
    double* p = x->arena + x->arena_used;
    x->arena_used += n;
    return p;
}
static pp_t* pp_alloc(nuts_ctx* x) {
    if (x->pp_used >= x->pp_cap) abort();
    pp_t* p = &x->pps[x->pp_used++];
    p->theta = arena_alloc(x, (size_t)x->D);
    p->r = arena_alloc(x, (size_t)x->D);
    p->g = arena_alloc(x, (size_t)x->D);
    return p;
}

/* LogExpFunctions.logaddexp (call sites trajectory.jl:192,198) */
static double logaddexp(double a, double b) {
    double delta = (a == b) ? 0.0 : fabs(a - b);
    return jl_max(a, b) + log1p(exp(-delta));
}
static double maxabs(double a, double b) { return fabs(a) > fabs(b) ? a : b; } /* trajectory.jl:526 */

static term_t term_mul(term_t a, term_t b) { /* trajectory.jl:491-493 */
    term_t t = {a.dynamic || b.dynamic, a.numerical || b.numerical};
    return t;
}
static int is_term(term_t t) { return t.dynamic || t.numerical; }

/* generalised_uturn_criterion(rho, p_sharp_minus, p_sharp_plus) (trajectory.jl:615-617) */
static int gen_crit(nuts_ctx* x, const double* rho, const double* rminus, const double* rplus) {
    double* a = x->scratch;
    double dl = 0.0, dr_ = 0.0;
    orc_dHdr(x->me, x->D, x->c, rminus, a);
    for (int d = 0; d < x->D; ++d) dl += rho[d] * a[d];
    orc_dHdr(x->me, x->D, x->c, rplus, a);
    for (int d = 0; d < x->D; ++d) dr_ += rho[d] * a[d];
    return (dl <= 0) || (dr_ <= 0);
}

/* isterminated(criterion, h, t, tleft, tright) (trajectory.jl:551-613) */
static term_t uturn(nuts_ctx* x, const tree_t* t, const tree_t* tl, const tree_t* tr) {
    term_t r = {0, 0};
    const int D = x->D;
    if (x->criterion == 1) { /* ClassicNoUTurn (trajectory.jl:551-557) */
        double* a = x->scratch;
        double* nr = x->scratch + D;
        double s1 = 0.0, s2 = 0.0;
        for (int d = 0; d < D; ++d) nr[d] = -t->zleft->r[d];
        orc_dHdr(x->me, D, x->c, nr, a); /* dH/dr(h, -z0.r) */
        for (int d = 0; d < D; ++d) s1 += (t->zright->theta[d] - t->zleft->theta[d]) * a[d];
        orc_dHdr(x->me, D, x->c, t->zright->r, a);
        for (int d = 0; d < D; ++d) s2 += (-(t->zright->theta[d] - t->zleft->theta[d])) * a[d];
        r.dynamic = (s1 >= 0) || (s2 >= 0);
        return r;
    }
    r.dynamic = gen_crit(x, t->rho, t->zleft->r, t->zright->r); /* :566-570 */
    if (x->criterion == 2) { /* StrictGeneralisedNoUTurn (:579-613) */
        double* rho = (double*)malloc(sizeof(double) * (size_t)D);
        for (int d = 0; d < D; ++d) rho[d] = tl->rho[d] + tr->zleft->r[d]; /* check_left_subtree */
        int s2 = gen_crit(x, rho, t->zleft->r, tr->zleft->r);
        for (int d = 0; d < D; ++d) rho[d] = tl->zright->r[d] + tr->rho[d]; /* check_right_subtree */
        int s3 = gen_crit(x, rho, tl->zright->r, t->zright->r);
        free(rho);
        r.dynamic = r.dynamic || s2 || s3;
    }
    return r;
}

static void build_tree(nuts_ctx* x, const pp_t* z, sampler_t sampler, int v, int j, double H0, tree_t* tree_out,
                       sampler_t* sampler_out, term_t* term_out) {
    const int D = x->D;
    if (j == 0) { /* trajectory.jl:638-647 */
        pp_t* z1 = pp_alloc(x);
        memcpy(z1->theta, z->theta, sizeof(double) * (size_t)D);
        memcpy(z1->r, z->r, sizeof(double) * (size_t)D);
        memcpy(z1->g, z->g, sizeof(double) * (size_t)D);
        double e = v > 0 ? x->eps : -x->eps; /* step(..., v): fwd = v>0 (integrator.jl:221-226) */
        lf_one_step(x->m, x->me, D, x->c, e, 1, 1, g_temper_alpha, z1->theta, z1->r, z1->g, &z1->lp, &z1->lk, x->scratch,
                    x->scratch + D);
        double H1 = -(z1->lp + z1->lk);
        double dH = H1 - H0;
        double a1 = exp(jl_min(0.0, -dH));
        tree_out->zleft = z1;
        tree_out->zright = z1;
        tree_out->rho = arena_alloc(x, (size_t)D);
        memcpy(tree_out->rho, z1->r, sizeof(double) * (size_t)D); /* TurnStatistic(z.r) :462-464 */
        tree_out->sum_alpha = a1;
        tree_out->n_alpha = 1;
        tree_out->dH_max = dH;
        sampler_out->zcand = z1;
        term_out->dynamic = 0;
        if (x->sampler == 1) { /* SliceTS(s, H0, zcand) :164-166; Termination(::SliceTS) :500-502 */
            sampler_out->lw = (x->lu <= (z1->lp + z1->lk)) ? 1.0 : 0.0;
            term_out->numerical = !(x->lu < x->delta_max + -H1);
        } else {
            sampler_out->lw = H0 + (z1->lp + z1->lk); /* MultinomialTS(s,H0,zcand) :174-176 */
            term_out->numerical = !(-H0 < x->delta_max + -H1); /* :503-507 */
        }
        (void)sampler;
        return;
    }
    tree_t t1;
    sampler_t s1;
    term_t e1;
    build_tree(x, z, sampler, v, j - 1, H0, &t1, &s1, &e1); /* :651 */
    if (!is_term(e1)) {                                    /* :653 */
        tree_t t2, tl, tr;
        sampler_t s2;
        term_t e2;
        if (v == -1) { /* :655-660 */
            build_tree(x, t1.zleft, sampler, v, j - 1, H0, &t2, &s2, &e2);
            tl = t2;
            tr = t1;
        } else { /* :661-665 */
            build_tree(x, t1.zright, sampler, v, j - 1, H0, &t2, &s2, &e2);
            tl = t1;
            tr = t2;
        }
        /* combine(treeleft, treeright) :533-542 */
        tree_t t;
        t.zleft = tl.zleft;
        t.zright = tr.zright;
        t.rho = arena_alloc(x, (size_t)D);
        for (int d = 0; d < D; ++d) t.rho[d] = tl.rho[d] + tr.rho[d];
        t.sum_alpha = tl.sum_alpha + tr.sum_alpha;
        t.n_alpha = tl.n_alpha + tr.n_alpha;
        t.dH_max = maxabs(tl.dH_max, tr.dH_max);
        sampler_t s;
        if (x->sampler == 1) { /* combine(rng, s1::SliceTS, s2) :178-183 */
            double n = s1.lw + s2.lw;
            double u = x->exps[x->n_exp++]; /* rand(rng) */
            s.zcand = (n * u < s1.lw) ? s1.zcand : s2.zcand;
            s.lw = n;
        } else { /* combine(rng, sampler', sampler'') :191-195 */
            double lw = logaddexp(s1.lw, s2.lw);
            double ex = x->exps[x->n_exp++];
            s.zcand = (lw < s1.lw + ex) ? s1.zcand : s2.zcand;
            s.lw = lw;
        }
        e1 = term_mul(term_mul(e1, e2), uturn(x, &t, &tl, &tr)); /* :668-671 */
        t1 = t;
        s1 = s;
    }
    *tree_out = t1;
    *sampler_out = s1;
    *term_out = e1;
}

void orc_nuts_transition(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                         const double* eps_chain, int32_t max_depth, double delta_max,
                         const double* normal_tape, const uint8_t* dir_tape, int64_t dir_stride,
                         const double* exp_tape, int64_t exp_stride, const orc_phasepoint* z_in,
                         const orc_phasepoint* z_out, const orc_stats* st, int32_t* exp_used) {
    orc_nuts_transition_ex(m, me, D, N, eps, eps_chain, max_depth, delta_max, 0, 0, normal_tape, dir_tape, dir_stride,
                           exp_tape, exp_stride, z_in, z_out, st, exp_used);
}

void orc_nuts_transition_ex(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                            const double* eps_chain, int32_t max_depth, double delta_max, int32_t sampler_kind,
                            int32_t criterion, const double* normal_tape, const uint8_t* dir_tape, int64_t dir_stride,
                            const double* exp_tape, int64_t exp_stride, const orc_phasepoint* z_in,
                            const orc_phasepoint* z_out, const orc_stats* st, int32_t* exp_used) {
    size_t max_leaves = ((size_t)1 << max_depth) + 4;
    nuts_ctx x;
    x.sampler = sampler_kind;
    x.criterion = criterion;
    x.lu = 0.0;
    x.m = m;
    x.me = me;
    x.D = D;
    x.delta_max = delta_max;
    x.arena_cap = (size_t)D * (max_leaves * 5 + 64);
    x.arena = (double*)malloc(sizeof(double) * x.arena_cap);
    x.pp_cap = max_leaves;
    x.pps = (pp_t*)malloc(sizeof(pp_t) * x.pp_cap);
    x.scratch = (double*)malloc(sizeof(double) * (size_t)D * 2);
    double* grad = (double*)malloc(sizeof(double) * (size_t)D);
    for (int64_t c = 0; c < N; ++c) {
        x.c = c;
        x.eps = eps_chain ? eps_chain[c] : eps;
        x.dirs = dir_tape + dir_stride * c;
        x.exps = exp_tape + exp_stride * c;
        x.n_exp = 0;
        x.arena_used = 0;
        x.pp_used = 0;
        /* refresh (sampler.jl:55; hamiltonian.jl:213-220) */
        pp_t* z0 = pp_alloc(&x);
        memcpy(z0->theta, z_in->theta + z_in->ld * c, sizeof(double) * (size_t)D);
        if (normal_tape)
            refresh_momentum(me, D, c, normal_tape + (size_t)D * c, z_in->r + z_in->ld * c, z0->r);
        else
            memcpy(z0->r, z_in->r + z_in->ld * c, sizeof(double) * (size_t)D);
        double lp;
        orc_logp_grad(m, z0->theta, &lp, grad);
        for (int d = 0; d < D; ++d) z0->g[d] = -grad[d];
        z0->lp = map_nonfinite(lp);
        z0->lk = map_nonfinite(orc_neg_kinetic(me, D, c, z0->r));
        /* transition (trajectory.jl:677-742) */
        double H0 = -(z0->lp + z0->lk);
        tree_t tree;
        tree.zleft = z0;
        tree.zright = z0;
        tree.rho = arena_alloc(&x, (size_t)D);
        memcpy(tree.rho, z0->r, sizeof(double) * (size_t)D);
        tree.sum_alpha = 0.0;
        tree.n_alpha = 0;
        tree.dH_max = 0.0;
        sampler_t sampler = {z0, 0.0}; /* MultinomialTS(rng, z0) :155 */
        if (x.sampler == 1) {          /* SliceTS(rng, z0) = SliceTS(z0, neg_energy(z0) - randexp(rng), 1) :144-145 */
            x.lu = (z0->lp + z0->lk) - x.exps[x.n_exp++];
            sampler.lw = 1.0;
        }
        term_t term = {0, 0};
        const pp_t* zcand = z0;
        int j = 0, ndir = 0;
        while (!is_term(term) && j < max_depth) { /* :691 */
            int vleft = x.dirs[ndir++];           /* :693 */
            tree_t t1, tl, tr;
            sampler_t s1;
            term_t e1;
            if (vleft) {
                build_tree(&x, tree.zleft, sampler, -1, j, H0, &t1, &s1, &e1);
                tl = t1;
                tr = tree;
            } else {
                build_tree(&x, tree.zright, sampler, 1, j, H0, &t1, &s1, &e1);
                tl = tree;
                tr = t1;
            }
            if (!is_term(e1)) { /* :708-713 */
                j = j + 1;
                double ex = x.exps[x.n_exp++];
                if (x.sampler == 1) {
                    if (sampler.lw * ex < s1.lw) zcand = s1.zcand; /* mh_accept(::SliceTS): s.n * rand < s'.n :202 */
                } else if (sampler.lw < s1.lw + ex) zcand = s1.zcand; /* mh_accept :204-206 */
            }
            tree_t t; /* :715 */
            t.zleft = tl.zleft;
            t.zright = tr.zright;
            t.rho = arena_alloc(&x, (size_t)D);
            for (int d = 0; d < D; ++d) t.rho[d] = tl.rho[d] + tr.rho[d];
            t.sum_alpha = tl.sum_alpha + tr.sum_alpha;
            t.n_alpha = tl.n_alpha + tr.n_alpha;
            t.dH_max = maxabs(tl.dH_max, tr.dH_max);
            tree = t;
            sampler.zcand = zcand; /* combine(zcand, sampler, sampler') :197-200, :717 */
            sampler.lw = (x.sampler == 1) ? sampler.lw + s1.lw : logaddexp(sampler.lw, s1.lw);
            term = term_mul(term_mul(term, e1), uturn(&x, &tree, &tl, &tr)); /* :719-722 */
        }
        double H = -(zcand->lp + zcand->lk);
        for (int d = 0; d < D; ++d) {
            z_out->theta[z_out->ld * c + d] = zcand->theta[d];
            z_out->r[z_out->ld * c + d] = zcand->r[d];
            z_out->lp_gradient[z_out->ld * c + d] = zcand->g[d];
        }
        z_out->lp_value[c] = zcand->lp;
        z_out->lk_value[c] = zcand->lk;
        if (z_out->lk_gradient) orc_dHdr(me, D, c, zcand->r, z_out->lk_gradient + z_out->ld * c);
        if (st) { /* :725-739 */
            if (st->n_steps) st->n_steps[c] = tree.n_alpha;
            if (st->is_accept) st->is_accept[c] = 1;
            if (st->acceptance_rate) st->acceptance_rate[c] = tree.sum_alpha / tree.n_alpha;
            if (st->log_density) st->log_density[c] = zcand->lp;
            if (st->hamiltonian_energy) st->hamiltonian_energy[c] = H;
            if (st->hamiltonian_energy_error) st->hamiltonian_energy_error[c] = H - H0;
            if (st->max_hamiltonian_energy_error) st->max_hamiltonian_energy_error[c] = tree.dH_max;
            if (st->tree_depth) st->tree_depth[c] = j;
            if (st->numerical_error) st->numerical_error[c] = (uint8_t)term.numerical;
        }
        if (exp_used) exp_used[c] = x.n_exp;
    }
    free(grad);
    free(x.scratch);
    free(x.pps);
    free(x.arena);
}

/* ------------------------------------------------------------------------------------------ */
/* adaptation                                                                                  */
/* ------------------------------------------------------------------------------------------ */
void orc_da_init(orc_da_state* s, int64_t n) { /* DAState(eps) stepsize.jl:27-36: mu = log(10 eps) */
    s->m = 0;
    for (int64_t i = 0; i < n; ++i) {
        s->mu[i] = log(10 * s->eps[i]);
        s->x_bar[i] = 0.0;
        s->H_bar[i] = 0.0;
    }
}
void orc_da_reset(orc_da_state* s, int64_t n) { orc_da_init(s, n); } /* stepsize.jl:38-52 */
void orc_da_finalize(orc_da_state* s, int64_t n) {                   /* stepsize.jl:54-62 */
    for (int64_t i = 0; i < n; ++i) s->eps[i] = exp(s->x_bar[i]);
}
void orc_da_adapt(orc_da_state* s, int64_t n, double gamma, double t0, double kappa, double delta,
                  const double* alpha) { /* stepsize.jl:178-210 */
    int64_t m = s->m + 1;
    double eta_H = 1.0 / ((double)m + t0);
    double eta_x = pow((double)m, -kappa);
    double sq = sqrt((double)m) / gamma;
    double* nx = (double*)malloc(sizeof(double) * (size_t)n * 3);
    double *nH = nx + n, *ne = nx + 2 * n;
    int all_fin = 1;
    for (int64_t i = 0; i < n; ++i) {
        double H_bar = (1.0 - eta_H) * s->H_bar[i] + eta_H * (delta - jl_min(1.0, alpha[i]));
        double xx = s->mu[i] - H_bar * sq;
        double x_bar = (1.0 - eta_x) * s->x_bar[i] + eta_x * xx;
        double e = exp(xx);
        nx[i] = x_bar;
        nH[i] = H_bar;
        ne[i] = e;
        if (!isfinite(e)) all_fin = 0;
    }
    if (all_fin) { /* else: revert everything incl. m (stepsize.jl:199-203) */
        s->m = m;
        for (int64_t i = 0; i < n; ++i) {
            s->x_bar[i] = nx[i];
            s->H_bar[i] = nH[i];
            s->eps[i] = ne[i];
        }
    }
    free(nx);
}

void orc_welford_var_push(int64_t* n, double* mu, double* M, int64_t len, const double* s) {
    *n += 1; /* massmatrix.jl:141-149 */
    double nn = (double)*n;
    for (int64_t i = 0; i < len; ++i) {
        double delta = s[i] - mu[i];
        mu[i] = mu[i] + delta / nn;
        M[i] = M[i] + delta * delta * ((nn - 1) / nn);
    }
}
void orc_welford_var_estimate(int64_t n, const double* M, int64_t len, double* var_out) {
    double nn = (double)n, e = 1e-3; /* massmatrix.jl:152-157 */
    for (int64_t i = 0; i < len; ++i) var_out[i] = nn / ((nn + 5) * (nn - 1)) * M[i] + e * (5 / (nn + 5));
}
void orc_welford_cov_push(int64_t* n, double* mu, double* M, int32_t D, const double* s) {
    *n += 1; /* massmatrix.jl:324-332 */
    double nn = (double)*n;
    double* delta = (double*)malloc(sizeof(double) * (size_t)D);
    for (int d = 0; d < D; ++d) {
        delta[d] = s[d] - mu[d];
        mu[d] = mu[d] + delta[d] / nn;
    }
    for (int k = 0; k < D; ++k)
        for (int d = 0; d < D; ++d) M[d + (int64_t)D * k] = M[d + (int64_t)D * k] + (s[d] - mu[d]) * delta[k];
    free(delta);
}
void orc_welford_cov_estimate(int64_t n, const double* M, int32_t D, double* cov_out) {
    double nn = (double)n, e = 1e-3; /* massmatrix.jl:335-340 */
    for (int k = 0; k < D; ++k)
        for (int d = 0; d < D; ++d) {
            double v = nn / ((nn + 5) * (nn - 1)) * M[d + (int64_t)D * k];
            if (d == k) v += e * (5 / (nn + 5));
            cov_out[d + (int64_t)D * k] = v;
        }
}

int32_t orc_stan_windows(int32_t init_buffer, int32_t term_buffer, int32_t window_size, int32_t n_adapts,
                         int32_t* window_start, int32_t* window_end, int32_t* splits_out) {
    /* stan_adaptor.jl:13-50 */
    int32_t ws = init_buffer + 1;
    int32_t we = n_adapts - term_buffer;
    int32_t ns = 0;
    int32_t next_window = init_buffer + window_size;
    while (next_window <= we) {
        int32_t next_window_boundary = next_window + 2 * window_size;
        if (next_window_boundary > we) next_window = we;
        if (ns < 64) splits_out[ns++] = next_window;
        window_size *= 2;
        next_window += window_size;
    }
    if (ns > 0 && splits_out[ns - 1] == n_adapts) ns--;
    *window_start = ws;
    *window_end = we;
    return ns;
}
