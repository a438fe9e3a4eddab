/*
 * ahmc_oracle.h -- CPU restatement of AdvancedHMC.jl's vectorised leapfrog / HMC / NUTS path.
 *
 * *** TEST INFRASTRUCTURE ONLY. ***  Nothing in the product (advancedhmc.jl_b200/, the C-ABI
 * library) may include, link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or the timed CPU baseline.
 *
 * PARITY PIN STATUS: the reference is pure Julia and no Julia binary exists in this image, and the
 * reference's tests hold no golden vectors for this path.  This restatement is
 * therefore pinned by (i) 50-digit closed-form linear-map known answers for Gaussian targets
 * (tests/golden/gen_closed_form.py), (ii) the algebraic identities the reference's own tests assert
 * (test/hamiltonian.jl:54-79, test/integrator.jl:17-32,108-153, test/adaptation.jl:131-151,
 * test/trajectory.jl:249-325), (iii) an independent op-for-op numpy twin (oracle/oracle_np.py) and (iv) for the
 * NUTS transitions (both trajectory samplers, all three termination criteria, numerical termination) a second,
 * independently written recursive restatement of src/trajectory.jl:626-742 evaluated in 50-digit arithmetic
 * (tests/golden/gen_nuts_mp.py -> tests/golden/nuts_mp50.json: identical trees, outputs to 1e-10), and likewise
 * for the static transitions and the adaptors (gen_hmc_mp.py, gen_adapt_mp.py).
 * It is NOT pinned against outputs of the reference itself: "parity unpinned" at that level.
 *
 * All `file:line` citations are relative to /root/reference/.
 * Layout: Julia column-major D x N -- element (d, chain c) at d + ld*c.
 */
#ifndef AHMC_ORACLE_H
#define AHMC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_MODEL_STD_NORMAL = 0, ORC_MODEL_DIAG_GAUSS = 1, ORC_MODEL_DENSE_GAUSS = 2, ORC_MODEL_FUNNEL = 3 };
enum { ORC_METRIC_UNIT = 0, ORC_METRIC_DIAG = 1, ORC_METRIC_DENSE = 2 };

/* Built-in targets (the reference takes an arbitrary closure, hamiltonian.jl:45-48; these are the
 * closures our tests/benchmarks plug in):
 *  STD_NORMAL : lp = c0 - sum(th^2)/2
 *  DIAG_GAUSS : p0 = m[D], p1 = s[D] (std devs); per test/common.jl:40-45
 *               lp = c0 + sum( -(abs2(m-th)/s^2)/2 ), grad = (m-th)/s^2
 *  DENSE_GAUSS: p0 = mu[D], p1 = P[DxD] (precision, column-major); lp = c0 - (th-mu)'P(th-mu)/2
 *  FUNNEL     : v=th[0]; lp = c0 - v^2/18 - sum_{i>=1}( th_i^2 exp(-v) + v )/2   
 */
typedef struct {
    int32_t kind;
    int32_t D;
    const double* p0;
    const double* p1;
    double c0;
} orc_model;

/* metric.jl:17-35 (Unit), :52-72 (Diag; Minv is D or D x N when chain_stride==D), :89-120 (Dense:
 * Minv D x D column-major, cholU = cholesky(Symmetric(Minv)).U column-major upper). */
typedef struct {
    int32_t kind;
    const double* Minv;
    int64_t chain_stride;
    const double* cholU;
} orc_metric;

/* PhasePoint (hamiltonian.jl:88-107): lp_gradient holds MINUS grad log pi (hamiltonian.jl:45-48). */
typedef struct {
    double* theta;
    double* r;
    double* lp_value;
    double* lp_gradient;
    double* lk_value;
    double* lk_gradient; /* may be NULL */
    int64_t ld;
} orc_phasepoint;

/* --- model / metric primitives ------------------------------------------------------------- */
/* user closure dlp/dth: returns lp and grad (PLUS gradient of log pi) for one chain */
void orc_logp_grad(const orc_model* m, const double* th, double* lp, double* grad);
/* dHdr (hamiltonian.jl:50-68) for chain c */
void orc_dHdr(const orc_metric* me, int32_t D, int64_t c, const double* r, double* out);
/* neg_energy kinetic part (hamiltonian.jl:155-184) for chain c */
double orc_neg_kinetic(const orc_metric* me, int32_t D, int64_t c, const double* r);
/* rand_momentum (metric.jl:290-320) from a tape of standard normals z[D]: writes r[D] */
void orc_rand_momentum(const orc_metric* me, int32_t D, int64_t c, const double* z, double* r);

/* phasepoint(h, th, r) (hamiltonian.jl:115-119): fills lp_*, lk_* from theta, r for all chains */
void orc_make_phasepoint(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, const orc_phasepoint* z);

/* step(lf, h, z, n_steps) (integrator.jl:216-265).  eps_chain==NULL -> scalar eps.  n_steps signed.
 * temper_alpha <= 0 -> no tempering (Leapfrog), else TemperedLeapfrog(eps, alpha) (:198-209).
 * compat_break_all != 0 mirrors the reference's matrix-mode quirk: the first step at which ANY
 * chain is non-finite stops every chain (hamiltonian.jl:141-142 + integrator.jl:252-258).
 * Otherwise each chain stops on its own (= the reference applied one chain at a time).
 * steps_done[c] = number of steps actually taken (the non-finite step included). status bit0 = non-finite.
 * z_out may alias z_in. */
void orc_leapfrog(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                  const double* eps_chain, int32_t n_steps, double temper_alpha, const orc_phasepoint* z_in,
                  const orc_phasepoint* z_out, uint32_t* status, int32_t* steps_done, int compat_break_all);

/* full_trajectory=Val(true) (integrator.jl:229,249-261): traj arrays hold |n_steps| phase points,
 * point i at offset i*step_stride (theta/r/lp_gradient/lk_gradient) and i*N (lp_value/lk_value). */
void orc_leapfrog_trajectory(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                             const double* eps_chain, int32_t n_steps, double temper_alpha,
                             const orc_phasepoint* z_in, const orc_phasepoint* traj, int64_t step_stride,
                             int32_t* steps_done, int compat_break_all);

/* per-chain statistics of one transition (trajectory.jl:286-298, :726-739) */
typedef struct {
    int32_t* n_steps;
    uint8_t* is_accept;
    double* acceptance_rate;
    double* log_density;
    double* hamiltonian_energy;
    double* hamiltonian_energy_error;
    double* max_hamiltonian_energy_error; /* NUTS only */
    int32_t* tree_depth;                  /* NUTS only */
    uint8_t* numerical_error;
} orc_stats;

/* static HMC transition, EndPointTS (sampler.jl:48-58 refresh + trajectory.jl:271-300,336-340,863-880).
 * normal_tape: D x N standard normals for the full momentum refresh (NULL -> keep z_in.r, no refresh);
 * exp_tape: N exponentials for mh_accept_ratio.  z_out gets the new phase point (momentum flipped). */
void orc_hmc_transition(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                        const double* eps_chain, int32_t n_steps, const double* normal_tape,
                        const double* exp_tape, const orc_phasepoint* z_in, const orc_phasepoint* z_out,
                        const orc_stats* st, int compat_break_all);
/* static HMC transition, MultinomialTS (trajectory.jl:344-390): n_steps_fwd forward + (n_steps - n_steps_fwd)
 * backward steps from z (the reference draws n_steps_fwd ONCE for all chains: `rand_coupled`, :371-373), the new
 * point is drawn from the whole trajectory with probabilities softmax(-H) by inverse-CDF on unif_tape[c]
 * (`randcat`, utilities.jl:92-103); acceptance_rate = mean_i min(1, exp(H0 - H_i)); is_accept = true. */
void orc_hmc_multinomial_transition(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                                    const double* eps_chain, int32_t n_steps, int32_t n_steps_fwd,
                                    const double* normal_tape, const double* unif_tape, const orc_phasepoint* z_in,
                                    const orc_phasepoint* z_out, const orc_stats* st);

/* PartialMomentumRefreshment(alpha) (hamiltonian.jl:222-254) for the transitions below: when non-zero, the
 * refreshed momentum is alpha*z_in.r + sqrt(1-alpha^2)*rand_momentum(tape).  Process-global test knob. */
void orc_set_partial_refresh(double alpha);
/* TemperedLeapfrog(eps, alpha) (integrator.jl:174-209) as the integrator of the transitions below; 0 = plain Leapfrog.
 * Process-global test knob. */
void orc_set_tempering(double alpha);

/* NUTS transition, MultinomialTS + GeneralisedNoUTurn (trajectory.jl:626-742), one chain at a time
 * (the reference has no vectorised NUTS).  Tapes per chain c: dir_tape[c*dir_stride + k] = k-th
 * rand(Bool) (`vleft`, :693); exp_tape[c*exp_stride + k] = k-th randexp in the reference's
 * consumption order (post-order combines :667,:191-195, then mh_accept :711,:204-206).
 * exp_used[c] (nullable) reports how many exponentials were consumed. */
void orc_nuts_transition(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                         const double* eps_chain, int32_t max_depth, double delta_max,
                         const double* normal_tape, const uint8_t* dir_tape, int64_t dir_stride,
                         const double* exp_tape, int64_t exp_stride, const orc_phasepoint* z_in,
                         const orc_phasepoint* z_out, const orc_stats* st, int32_t* exp_used);

/* NUTS variants: sampler_kind 0 = MultinomialTS, 1 = SliceTS (trajectory.jl:102-206); criterion 0 = GeneralisedNoUTurn,
 * 1 = ClassicNoUTurn, 2 = StrictGeneralisedNoUTurn (trajectory.jl:414-452, 551-613).  With SliceTS the tape entries
 * are consumed as: [0] = the randexp of the slice variable (:144-145), then one UNIFORM per combine / mh_accept. */
void orc_nuts_transition_ex(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                            const double* eps_chain, int32_t max_depth, double delta_max, int32_t sampler_kind,
                            int32_t criterion, const double* normal_tape, const uint8_t* dir_tape, int64_t dir_stride,
                            const double* exp_tape, int64_t exp_stride, const orc_phasepoint* z_in,
                            const orc_phasepoint* z_out, const orc_stats* st, int32_t* exp_used);

/* --- adaptation (src/adaptation) ----------------------------------------------------------- */
/* NesterovDualAveraging state (stepsize.jl:13-62) for n independent entries (scalar: n=1) */
typedef struct {
    int64_t m;
    double* eps;
    double* mu;
    double* x_bar;
    double* H_bar;
} orc_da_state;
void orc_da_init(orc_da_state* s, int64_t n);                       /* stepsize.jl:25-36 (eps preset) */
void orc_da_adapt(orc_da_state* s, int64_t n, double gamma, double t0, double kappa, double delta,
                  const double* alpha);                             /* stepsize.jl:178-210 */
void orc_da_reset(orc_da_state* s, int64_t n);                      /* stepsize.jl:38-52 */
void orc_da_finalize(orc_da_state* s, int64_t n);                   /* stepsize.jl:54-62 */

/* WelfordVar push (massmatrix.jl:141-149) over `len` independent entries; *n incremented */
void orc_welford_var_push(int64_t* n, double* mu, double* M, int64_t len, const double* s);
/* get_estimation (massmatrix.jl:152-157) */
void orc_welford_var_estimate(int64_t n, const double* M, int64_t len, double* var_out);
/* WelfordCov push / estimate (massmatrix.jl:324-340); M is D x D column-major */
void orc_welford_cov_push(int64_t* n, double* mu, double* M, int32_t D, const double* s);
void orc_welford_cov_estimate(int64_t n, const double* M, int32_t D, double* cov_out);

/* Stan window schedule (stan_adaptor.jl:13-50). splits_out must hold >= 64 entries.
 * returns number of splits; writes window_start/window_end. */
int32_t orc_stan_windows(int32_t init_buffer, int32_t term_buffer, int32_t window_size, int32_t n_adapts,
                         int32_t* window_start, int32_t* window_end, int32_t* splits_out);

/* multi-thread variant used ONLY as the timed "good CPU implementation" baseline (bench.py):
 * same arithmetic as orc_leapfrog, chains split across OpenMP threads. */
void orc_leapfrog_omp(const orc_model* m, const orc_metric* me, int32_t D, int64_t N, double eps,
                      const double* eps_chain, int32_t n_steps, const orc_phasepoint* z_in,
                      const orc_phasepoint* z_out, int32_t n_threads);

#ifdef __cplusplus
}
#endif
#endif
