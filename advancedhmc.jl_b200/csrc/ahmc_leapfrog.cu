// ahmc_leapfrog.cu -- K1: the fused leapfrog trajectory kernel (`step(lf, h, z, n_steps)`,
// src/integrator.jl:216-265) + `phasepoint` (src/hamiltonian.jl:115-119) + `rand_momentum`
// (src/metric.jl:290-320).
//
// One launch integrates every chain for all n_steps: a chain's theta, r and -grad(lp) are read once
// from HBM (coalesced runs of G doubles per group), stay in registers for the whole trajectory and
// are written once.  Two code paths share the kernel:
//
//  * EXACT path (every model x metric): per step the reference's op sequence with FMAs, the kinetic
//    and potential energies reduced by warp shuffles, and the reference's `isfinite(z)` test
//    (hamiltonian.jl:141-142) -- a non-finite chain stops on its own (default) and its phase point at
//    the break step is what is returned, as integrator.jl:252-258 does.
//
//  * FAST path (separable Gaussian targets STD_NORMAL / DIAG_GAUSS with Unit / Diag metric, no
//    tempering): state is kept in shifted coordinates x = theta - m, the two half kicks of
//    consecutive steps are merged and the per-coordinate constants a = eps*Minv, b = eps/s^2 are
//    precomputed, so a step costs 2 dependent DFMAs per coordinate and no reduction.  Exactness of the
//    reference's per-step `isfinite` control flow is kept by a magnitude argument: with
//    K = (1+max|a|)(1+max|b|) the sup-norm of (x, r) grows by at most K per step, so checking the
//    exponent fields of (x, r) against 2^200 every floor(100/log2 K) steps PROVES every intermediate
//    phase point (incl. its energies) was finite; a chain that fails a check (or whose parameters are
//    outside the proof's range) is simply re-run by the exact path inside the same launch.
#ifndef __CUDACC_RTC__
#include <cstdlib>
#endif
#include "ahmc_kernels.cuh"
#include "ahmc_traj.cuh"

namespace ahmc {

bool pick_layout(int D, int* G, int* E) {
    if (D < 1) return false;
    if (D <= 4) { *G = 4; *E = 1; return true; }
    if (D <= 8) { *G = 8; *E = 1; return true; }
    if (D <= 16) { *G = 16; *E = 1; return true; }
    if (D <= 32) { *G = 32; *E = 1; return true; }
    if (D <= 64) { *G = 32; *E = 2; return true; }
    if (D <= 128) { *G = 32; *E = 4; return true; }
    if (D <= 256) { *G = 32; *E = 8; return true; }
    if (D <= 512) { *G = 32; *E = 16; return true; }
    return false;
}

// K1 functor: start state = z_in, end state -> z_out (+ status / steps_done / min_break).  CONTIG: the fast path moves
// the state with lane-contiguous 128/256-bit accesses (init_c / done_c); the exact path always uses the interleaved form.
template <int G, int E, bool CONTIG = false>
struct StepIO {
    static constexpr bool kContig = CONTIG;
    const LeapfrogArgs& a;
    long long chain;
    int l;
    __device__ __forceinline__ bool has_g() const { return a.g_in != nullptr; }
    __device__ __forceinline__ void init(double (&th)[E], double (&r)[E], double (&g)[E]) const {
        vload_nc<G, E>(th, a.th_in + a.ld_in * chain, l, a.D);
        vload_nc<G, E>(r, a.r_in + a.ld_in * chain, l, a.D);
        if (a.g_in) vload_nc<G, E>(g, a.g_in + a.ld_in * chain, l, a.D);
    }
    __device__ __forceinline__ void init_c(double (&th)[E], double (&r)[E], double (&g)[E]) const {
        cload<E>(th, a.th_in + a.ld_in * chain, l, a.D);
        cload<E>(r, a.r_in + a.ld_in * chain, l, a.D);
        // issued unconditionally so that all three loads are in flight together (without a cached gradient theta is
        // read a second time and the value ignored)
        cload<E>(g, (a.g_in ? a.g_in : a.th_in) + a.ld_in * chain, l, a.D);
    }
    __device__ __forceinline__ void scalars(double lp, double lk, bool fin, int steps) const {
        if (l == 0) {
            a.lp_out[chain] = lp;
            a.lk_out[chain] = lk;
            if (a.status) a.status[chain] = fin ? 0u : AHMC_STATUS_NONFINITE;
            if (a.steps_done) a.steps_done[chain] = steps;
            if (!fin && a.min_break) atomicMin(a.min_break, steps);
        }
    }
    __device__ __forceinline__ void done(const double (&th)[E], const double (&r)[E], const double (&g)[E],
                                         const double (&dr)[E], double lp, double lk, bool fin, int steps) const {
        vstore<G, E>(a.th_out + a.ld_out * chain, th, l, a.D);
        vstore<G, E>(a.r_out + a.ld_out * chain, r, l, a.D);
        vstore<G, E>(a.g_out + a.ld_out * chain, g, l, a.D);
        if (a.dr_out) vstore<G, E>(a.dr_out + a.ld_out * chain, dr, l, a.D);
        scalars(lp, lk, fin, steps);
    }
    __device__ __forceinline__ void done_c(const double (&th)[E], const double (&r)[E], const double (&g)[E],
                                           const double (&dr)[E], double lp, double lk, bool fin, int steps) const {
        cstore<E>(a.th_out + a.ld_out * chain, th, l, a.D);
        cstore<E>(a.r_out + a.ld_out * chain, r, l, a.D);
        cstore<E>(a.g_out + a.ld_out * chain, g, l, a.D);
        if (a.dr_out) cstore<E>(a.dr_out + a.ld_out * chain, dr, l, a.D);
        scalars(lp, lk, fin, steps);
    }
};

// Occupancy: the headline batch (4096 chains x D=128 -> 4096 warps) fits the chip in ONE wave only if 7 blocks of
// 4 warps are resident per SM (148 x 7 x 4 = 4144 warp slots), i.e. <= 72 registers per thread.  The fast path needs
// ~50; the cap makes the (rarely taken) exact fallback of the separable kernels spill a little, which is the right
// trade.  Wider layouts (E >= 8) and non-separable models keep the default budget.
template <int MODEL, int METRIC, int E>
constexpr int min_blocks_per_sm() {
    return (FastCapable<MODEL, METRIC>::value && E <= 4) ? 7 : 1;
}
// The funnel trajectory kernel (exact path only) sits at 95 registers uncapped = 5 blocks per SM = two waves for 4096
// chains; capped it spills ~50 bytes outside the step loop and runs in one.  Its transition kernel would spill inside
// the loop, so that one keeps the default.
// The transition kernel of a general target: 128 registers (4 blocks per SM) are enough for E <= 4 without spills; left
// alone it takes ~140 and loses a block per SM (measured on the funnel: 64 us per transition vs 53).
template <int MODEL, int METRIC, int E>
constexpr int min_blocks_hmc() {
    return (FastCapable<MODEL, METRIC>::value && E <= 4) ? 7 : ((METRIC != AHMC_METRIC_DENSE && MODEL != AHMC_MODEL_DENSE_GAUSS && E <= 4) ? 4 : 1);
}
template <int MODEL, int METRIC, int E>
constexpr int min_blocks_lf() {
    return (MODEL == AHMC_MODEL_FUNNEL && METRIC != AHMC_METRIC_DENSE && E <= 4) ? 7 : min_blocks_per_sm<MODEL, METRIC, E>();
}

template <int MODEL, int METRIC, int G, int E, bool CONTIG = false>
__global__ void __launch_bounds__(kBlockThreads, min_blocks_lf<MODEL, METRIC, E>()) leapfrog_kernel(const LeapfrogArgs a) {
    extern __shared__ double smem[];
    const int l = threadIdx.x % G;
    const int grp_in_block = threadIdx.x / G;
    const long long chain0 = (long long)blockIdx.x * (kBlockThreads / G) + grp_in_block;
    const long long chain = chain0 < a.N ? chain0 : a.N - 1;  // tail groups shadow the last chain, never store
    const bool valid = chain0 < a.N && (!a.only_mask || a.only_mask[chain] != 0);
    double* xs = smem + (size_t)grp_in_block * slab_vectors<MODEL>() * a.D;
    double eps = a.eps_chain ? __ldg(a.eps_chain + chain) : a.eps;
    eps = a.fwd ? eps : -eps;  // integrator.jl:226
    StepIO<G, E, CONTIG> io{a, chain, l};
    run_trajectory<MODEL, METRIC, G, E>(a.model, a.metric, a.D, chain, valid, l, xs, eps, a.n_steps, a.temper_alpha,
                                        a.flags, io);
}

// ---------------------------------------------------------------------------------------------
// K2: one static-HMC transition (sampler.jl:48-58 + trajectory.jl:271-300, 312-340, 863-880)
// ---------------------------------------------------------------------------------------------
template <int METRIC, int G, int E>
struct HmcIO {
    const HmcArgs& h;
    long long chain;
    int l;
    const double* src_th;  // start point of THIS transition (z_in for the first, z_out afterwards)
    const double* src_g;
    long long stat_idx;    // t*N + chain
    double H0, lp0, lk0, ex;
    double r0[E];

    static constexpr bool kContig = false;
    __device__ __forceinline__ bool has_g() const { return true; }
    __device__ __forceinline__ void init(double (&th)[E], double (&r)[E], double (&g)[E]) const {
        vload_nc<G, E>(th, src_th, l, h.lf.D);
        vload_nc<G, E>(g, src_g, l, h.lf.D);
#pragma unroll
        for (int e = 0; e < E; ++e) r[e] = r0[e];
    }
    // mh_accept_ratio + accept_phasepoint! + momentum flip + stats
    __device__ __forceinline__ void done(const double (&th)[E], const double (&r)[E], const double (&g)[E],
                                         const double (&dr)[E], double lp, double lk, bool fin, int steps) const {
        const LeapfrogArgs& a = h.lf;
        const double H1 = -(lp + lk);                   // energy(z') (hamiltonian.jl:149,194)
        const bool accept = H1 < H0 + ex;               // trajectory.jl:869-877
        double alpha = exp(H0 - H1);                    // min(1, exp(H - H')) with Julia's NaN-propagating min
        alpha = (alpha != alpha) ? alpha : (alpha < 1.0 ? alpha : 1.0);
        double* tho = a.th_out + a.ld_out * chain;
        double* ro = a.r_out + a.ld_out * chain;
        double* go = a.g_out + a.ld_out * chain;
        double* dro = h.draws ? h.draws + stat_idx * a.D : nullptr;
        double lpn, lkn;
        if (accept) {
            double nr[E];
#pragma unroll
            for (int e = 0; e < E; ++e) nr[e] = -r[e];  // flip (trajectory.jl:283)
            vstore<G, E>(tho, th, l, a.D);
            vstore<G, E>(ro, nr, l, a.D);
            vstore<G, E>(go, g, l, a.D);
            if (dro) vstore<G, E>(dro, th, l, a.D);
            lpn = lp;
            lkn = lk;
        } else {  // revert (trajectory.jl:312-332)
            double t0[E], g0[E], nr[E];
            vload_nc<G, E>(t0, src_th, l, a.D);
            vload_nc<G, E>(g0, src_g, l, a.D);
#pragma unroll
            for (int e = 0; e < E; ++e) nr[e] = -r0[e];
            vstore<G, E>(tho, t0, l, a.D);
            vstore<G, E>(ro, nr, l, a.D);
            vstore<G, E>(go, g0, l, a.D);
            if (dro) vstore<G, E>(dro, t0, l, a.D);
            lpn = lp0;
            lkn = lk0;
        }
        if (l == 0) {
            const double H = -(lpn + lkn);
            a.lp_out[chain] = lpn;
            a.lk_out[chain] = lkn;
            if (a.status) a.status[chain] = fin ? 0u : AHMC_STATUS_NONFINITE;
            if (a.steps_done) a.steps_done[chain] = steps;
            const StatsDev& st = h.st;
            if (st.n_steps) st.n_steps[stat_idx] = a.n_steps;  // nsteps(tau), nominal (trajectory.jl:288)
            if (st.is_accept) st.is_accept[stat_idx] = accept ? 1 : 0;
            if (st.acceptance_rate) st.acceptance_rate[stat_idx] = alpha;
            if (st.log_density) st.log_density[stat_idx] = lpn;
            if (st.hamiltonian_energy) st.hamiltonian_energy[stat_idx] = H;
            if (st.hamiltonian_energy_error) st.hamiltonian_energy_error[stat_idx] = H - H0;
            if (st.numerical_error) st.numerical_error[stat_idx] = finite_d(H1) ? 0 : 1;
        }
        (void)dr;
    }
};

// One launch = n_transitions static-HMC transitions per chain (the reference's `for i in 1:n_samples` loop,
// sampler.jl:182, without adaptation): state is re-read from the output phase point, which stays L2-resident.
template <int MODEL, int METRIC, int G, int E>
__global__ void __launch_bounds__(kBlockThreads, min_blocks_hmc<MODEL, METRIC, E>()) hmc_kernel(const HmcArgs h) {
    extern __shared__ double smem[];
    const LeapfrogArgs& a = h.lf;
    const int l = threadIdx.x % G;
    const int grp_in_block = threadIdx.x / G;
    const long long chain0 = (long long)blockIdx.x * (kBlockThreads / G) + grp_in_block;
    const bool valid = chain0 < a.N;
    const long long chain = valid ? chain0 : a.N - 1;
    const int D = a.D;
    double* xs = smem + (size_t)grp_in_block * slab_vectors<MODEL>() * D;
    const double eps = a.eps_chain ? __ldg(a.eps_chain + chain) : a.eps;

    MetricOps<METRIC, G, E> me;
    me.load(a.metric, chain, l, D);
    HmcIO<METRIC, G, E> io{h, chain, l};
    for (int t = 0; t < h.n_transitions; ++t) {
        const bool first = (t == 0);
        io.src_th = first ? a.th_in + a.ld_in * chain : a.th_out + a.ld_out * chain;
        io.src_g = first ? a.g_in + a.ld_in * chain : a.g_out + a.ld_out * chain;
        io.stat_idx = (long long)t * a.N + chain;
        const uint64_t off = h.rng.offset + (uint64_t)t;
        // refresh (hamiltonian.jl:213-220): new momentum, kinetic energy; lp is the cached value (quirk Q2:
        // the reference recomputes it from theta -- same number)
        if (h.refresh) {
            if (h.rng.normal_tape) {
                vload_nc<G, E>(io.r0, h.rng.normal_tape + (long long)D * chain, l, D);
            } else {
                philox_normals<G, E>(h.rng.seed, off, chain, l, D, io.r0);
            }
            me.rand_momentum(io.r0, l);
            if (h.rng.partial_alpha != 0.0) {  // PartialMomentumRefreshment (hamiltonian.jl:243-254)
                double rp[E];
                vload_nc<G, E>(rp, first ? a.r_in + a.ld_in * chain : a.r_out + a.ld_out * chain, l, D);
                const double al = h.rng.partial_alpha, be = sqrt(1.0 - al * al);
#pragma unroll
                for (int e = 0; e < E; ++e) io.r0[e] = al * rp[e] + be * io.r0[e];
            }
        } else {
            vload_nc<G, E>(io.r0, first ? a.r_in + a.ld_in * chain : a.r_out + a.ld_out * chain, l, D);
        }
        {
            double dr0[E];
            io.lk0 = map_nonfinite(kinetic<METRIC, G, E>(me, io.r0, dr0, xs, l));
        }
        io.lp0 = map_nonfinite(first ? a.lp_in[chain] : a.lp_out[chain]);
        io.H0 = -(io.lp0 + io.lk0);
        io.ex = h.rng.exp_tape ? h.rng.exp_tape[chain] : philox_exp(h.rng.seed, off, chain, 0);
        run_trajectory<MODEL, METRIC, G, E>(a.model, a.metric, D, chain, valid, l, xs, eps, a.n_steps, h.rng.temper_alpha, a.flags, io);
        __syncwarp();
    }
}

// ---------------------------------------------------------------------------------------------
// find_good_stepsize (trajectory.jl:768-837), one independent search per chain, the WHOLE search in one launch:
// momentum draw, H, the direction probe, the crossing loop (doubling / halving until the one-step acceptance ratio
// crosses 1/2) and the bisection (until it lies in (1/4, 3/4]).  Every probe A(h, z, eps) (:753-757) is one exact
// leapfrog step from the chain's start point held in registers.  Loops are warp-uniform (groups that finished keep
// stepping with their final eps and ignore the result), so shuffles stay convergent when several chains share a warp.
// Mirrors the reference's control flow literally, including its quirk of probing with eps (not eps') in the crossing loop.
// ---------------------------------------------------------------------------------------------
template <int MODEL, int METRIC, int G, int E>
__global__ void __launch_bounds__(kBlockThreads) find_eps_kernel(const FindEpsArgs a) {
    extern __shared__ double smem[];
    const int l = threadIdx.x % G;
    const int grp_in_block = threadIdx.x / G;
    const long long chain0 = (long long)blockIdx.x * (kBlockThreads / G) + grp_in_block;
    const bool valid = chain0 < a.N;
    const long long chain = valid ? chain0 : a.N - 1;
    const int D = a.D;
    double* xs = smem + (size_t)grp_in_block * slab_vectors<MODEL>() * D;
    ModelOps<MODEL, G, E> mo;
    MetricOps<METRIC, G, E> me;
    mo.load(a.model, l, D);
    me.load(a.metric, chain, l, D);
    ChainState<E> z0;
    vload_nc<G, E>(z0.th, a.th + a.ld * chain, l, D);
    vload_nc<G, E>(z0.g, a.g + a.ld * chain, l, D);
    if (a.normal_tape) vload_nc<G, E>(z0.r, a.normal_tape + (long long)D * chain, l, D);
    else philox_normals<G, E>(a.seed, a.offset, chain, l, D, z0.r);
    me.rand_momentum(z0.r, l);
    if (a.r_out && valid) vstore<G, E>(a.r_out + a.ld * chain, z0.r, l, D);
    double dr[E];
    const double lk0 = map_nonfinite(kinetic<METRIC, G, E>(me, z0.r, dr, xs, l));
    const double H = -(map_nonfinite(a.lp[chain]) + lk0);  // energy(z) (hamiltonian.jl:149,194)
    auto probe = [&](double eps) -> double {                // H' of A(h, z, eps) (trajectory.jl:753-757)
        ChainState<E> s = z0;
        leapfrog_step<MODEL, METRIC, G, E>(s, mo, me, eps, dr, xs, l);
        return -(s.lp + s.lk);
    };
    const double log_a_min = 2.0 * -0.6931471805599453, log_a_cross = -0.6931471805599453, log_a_max = log(0.75);
    double eps = a.eps0, eps_p = a.eps0;
    double dH = H - probe(eps);
    const bool too_high = dH > log_a_cross;
    bool active = true;
    for (int it = 0; it < a.max_iters; ++it) {  // crossing step (:796-810)
        if (!__any_sync(FULL, active)) break;
        if (active) eps_p = too_high ? 2.0 * eps : 0.5 * eps;
        dH = H - probe(eps);
        if (active) {
            if (too_high != (dH > log_a_cross)) active = false;
            else eps = eps_p;
        }
    }
    double lo = fmin(eps, eps_p), hi = fmax(eps, eps_p);  // minmax (:818)
    active = true;
    for (int it = 0; it < a.max_iters; ++it) {  // bisection (:822-834)
        if (!__any_sync(FULL, active)) break;
        const double mid = 0.5 * (lo + hi);
        dH = H - probe(active ? mid : lo);
        if (active) {
            if (dH > log_a_max) lo = mid;
            else if (dH < log_a_min) hi = mid;
            else {
                lo = mid;
                active = false;
            }
        }
    }
    if (valid && l == 0) a.eps_out[chain] = lo;
}

// ---------------------------------------------------------------------------------------------
// phasepoint(h, theta, r)  (hamiltonian.jl:115-119)
// ---------------------------------------------------------------------------------------------
template <int MODEL, int METRIC, int G, int E>
__global__ void __launch_bounds__(kBlockThreads) phasepoint_kernel(const PhasepointArgs a) {
    extern __shared__ double smem[];
    const int l = threadIdx.x % G;
    const int grp_in_block = threadIdx.x / G;
    const long long chain0 = (long long)blockIdx.x * (kBlockThreads / G) + grp_in_block;
    const bool valid = chain0 < a.N;
    const long long chain = valid ? chain0 : a.N - 1;
    const int D = a.D;
    double* xs = smem + (size_t)grp_in_block * slab_vectors<MODEL>() * D;
    ModelOps<MODEL, G, E> mo;
    MetricOps<METRIC, G, E> me;
    mo.load(a.model, l, D);
    me.load(a.metric, chain, l, D);
    double th[E], r[E], g[E], dr[E];
    vload_nc<G, E>(th, a.th + a.ld * chain, l, D);
    vload_nc<G, E>(r, a.r + a.ld * chain, l, D);
    double lp = map_nonfinite(mo.eval(th, g, xs, l));
    double lk = map_nonfinite(kinetic<METRIC, G, E>(me, r, dr, xs, l));
    if (valid) {
        vstore<G, E>(a.g + a.ld * chain, g, l, D);
        if (a.dr) vstore<G, E>(a.dr + a.ld * chain, dr, l, D);
        if (l == 0) {
            a.lp[chain] = lp;
            a.lk[chain] = lk;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// rand_momentum  (metric.jl:290-320)
// ---------------------------------------------------------------------------------------------
template <int METRIC, int G, int E>
__global__ void __launch_bounds__(kBlockThreads) momentum_kernel(const MomentumArgs a) {
    const int l = threadIdx.x % G;
    const long long chain0 = (long long)blockIdx.x * (kBlockThreads / G) + threadIdx.x / G;
    const bool valid = chain0 < a.N;
    const long long chain = valid ? chain0 : a.N - 1;
    const int D = a.D;
    MetricOps<METRIC, G, E> me;
    me.load(a.metric, chain, l, D);
    double r[E];
    if (a.normal_tape) {
        vload_nc<G, E>(r, a.normal_tape + (long long)D * chain, l, D);
    } else {
        philox_normals<G, E>(a.seed, a.offset, chain, l, D, r);
    }
    me.rand_momentum(r, l);
    if (valid) vstore<G, E>(a.r + a.ld * chain, r, l, D);
}

// ---------------------------------------------------------------------------------------------
// split-step kernels: the generic path for a user-supplied gradient (hamiltonian.jl:45-48 closure)
// ---------------------------------------------------------------------------------------------
// first half of integrator.jl:235-240: temper, r -= eps/2 g, theta += eps dH/dr(r)
template <int METRIC, int G, int E>
__global__ void __launch_bounds__(kBlockThreads) kick_drift_kernel(const SplitArgs a) {
    extern __shared__ double smem[];
    const int l = threadIdx.x % G;
    const int grp_in_block = threadIdx.x / G;
    const long long chain0 = (long long)blockIdx.x * (kBlockThreads / G) + grp_in_block;
    const bool valid = chain0 < a.N;
    const long long chain = valid ? chain0 : a.N - 1;
    const int D = a.D;
    double* xs = smem + (size_t)grp_in_block * D;
    const bool active = valid && a.status[chain] == 0u;
    double eps = a.eps_chain ? __ldg(a.eps_chain + chain) : a.eps;
    eps = a.fwd ? eps : -eps;
    MetricOps<METRIC, G, E> me;
    me.load(a.metric, chain, l, D);
    double th[E], r[E], g[E], dr[E];
    vload_nc<G, E>(th, a.th + a.ld * chain, l, D);
    vload_nc<G, E>(r, a.r + a.ld * chain, l, D);
    vload_nc<G, E>(g, a.g + a.ld * chain, l, D);
    const double he = 0.5 * eps;
#pragma unroll
    for (int e = 0; e < E; ++e) r[e] = fma(-he, g[e], r[e] * a.mul);
    me.dHdr(r, dr, xs, l);
#pragma unroll
    for (int e = 0; e < E; ++e) th[e] = fma(eps, dr[e], th[e]);
    if (active) {
        vstore<G, E>(a.th + a.ld * chain, th, l, D);
        vstore<G, E>(a.r + a.ld * chain, r, l, D);
    }
}

// second half of integrator.jl:242-258: g = -grad, r -= eps/2 g, temper, energies, isfinite(z).
// no_kick = 1 turns it into the tail of `phasepoint` (hamiltonian.jl:115-119): r untouched, status ignored,
// cb_grad / cb_lp may be NULL (then only lk / dH/dr are produced).
template <int METRIC, int G, int E>
__global__ void __launch_bounds__(kBlockThreads) kick_energy_kernel(const SplitArgs a) {
    extern __shared__ double smem[];
    const int l = threadIdx.x % G;
    const int grp_in_block = threadIdx.x / G;
    const long long chain0 = (long long)blockIdx.x * (kBlockThreads / G) + grp_in_block;
    const bool valid = chain0 < a.N;
    const long long chain = valid ? chain0 : a.N - 1;
    const int D = a.D;
    double* xs = smem + (size_t)grp_in_block * D;
    const bool active = valid && (a.no_kick || a.status[chain] == 0u);
    double eps = a.eps_chain ? __ldg(a.eps_chain + chain) : a.eps;
    eps = a.fwd ? eps : -eps;
    MetricOps<METRIC, G, E> me;
    me.load(a.metric, chain, l, D);
    double r[E], g[E], dr[E];
    vload_nc<G, E>(r, a.r + a.ld * chain, l, D);
    if (a.cb_grad) {
        vload_nc<G, E>(g, a.cb_grad + a.ld * chain, l, D);
#pragma unroll
        for (int e = 0; e < E; ++e) g[e] = -g[e];  // dH/dtheta = DualValue(lp, -grad) (hamiltonian.jl:47)
    } else {
#pragma unroll
        for (int e = 0; e < E; ++e) g[e] = 0.0;
    }
    if (!a.no_kick) {
        const double he = 0.5 * eps;
#pragma unroll
        for (int e = 0; e < E; ++e) r[e] = fma(-he, g[e], r[e]) * a.mul;
    }
    const double lk = kinetic<METRIC, G, E>(me, r, dr, xs, l);
    const double lp = a.cb_lp ? a.cb_lp[chain] : 0.0;
    bool fin = true;
#pragma unroll
    for (int e = 0; e < E; ++e) fin = fin && finite_d(g[e]) && finite_d(dr[e]);
    fin = Grp<G>::all(fin) && finite_d(lp) && finite_d(lk);
    if (active) {
        if (!a.no_kick) vstore<G, E>(a.r + a.ld * chain, r, l, D);
        if (a.cb_grad) vstore<G, E>(a.g + a.ld * chain, g, l, D);
        if (a.dr) vstore<G, E>(a.dr + a.ld * chain, dr, l, D);
        if (l == 0) {
            if (a.cb_lp) a.lp[chain] = map_nonfinite(lp);
            a.lk[chain] = map_nonfinite(lk);
            if (!a.no_kick) {
                if (a.steps_done) a.steps_done[chain] = a.step_index;
                if (!fin) {
                    a.status[chain] = AHMC_STATUS_NONFINITE;
                    if (a.any_nonfinite) *a.any_nonfinite = 1;
                }
            }
        }
    }
}

// mh_accept_ratio + accept_phasepoint! + flip + stats, element-parallel over the group (trajectory.jl:271-300)
template <int G, int E>
__global__ void __launch_bounds__(kBlockThreads) mh_select_kernel(const MhArgs a) {
    const int l = threadIdx.x % G;
    const long long chain0 = (long long)blockIdx.x * (kBlockThreads / G) + threadIdx.x / G;
    if (chain0 >= a.N) return;
    const long long chain = chain0;
    const int D = a.D;
    const double lp0 = map_nonfinite(a.lp0[chain]), lk0 = a.lk0[chain];
    const double lp1 = a.lp[chain], lk1 = a.lk[chain];
    const double H0 = -(lp0 + lk0), H1 = -(lp1 + lk1);
    const double ex = a.rng.exp_tape ? a.rng.exp_tape[chain] : philox_exp(a.rng.seed, a.rng.offset, chain, 0);
    const bool accept = H1 < H0 + ex;
    double alpha = exp(H0 - H1);
    alpha = (alpha != alpha) ? alpha : (alpha < 1.0 ? alpha : 1.0);
    double t[E];
    if (accept) {
        vload_nc<G, E>(t, a.r + a.ld * chain, l, D);
#pragma unroll
        for (int e = 0; e < E; ++e) t[e] = -t[e];
        vstore<G, E>(a.r + a.ld * chain, t, l, D);
    } else {
        vload_nc<G, E>(t, a.th0 + a.ld0 * chain, l, D);
        vstore<G, E>(a.th + a.ld * chain, t, l, D);
        vload_nc<G, E>(t, a.g0 + a.ld0 * chain, l, D);
        vstore<G, E>(a.g + a.ld * chain, t, l, D);
        vload_nc<G, E>(t, a.r0 + (long long)D * chain, l, D);
#pragma unroll
        for (int e = 0; e < E; ++e) t[e] = -t[e];
        vstore<G, E>(a.r + a.ld * chain, t, l, D);
    }
    if (l == 0) {
        const double lpn = accept ? lp1 : lp0, lkn = accept ? lk1 : lk0;
        const double H = -(lpn + lkn);
        a.lp[chain] = lpn;
        a.lk[chain] = lkn;
        const StatsDev& st = a.st;
        if (st.n_steps) st.n_steps[chain] = a.n_steps;
        if (st.is_accept) st.is_accept[chain] = accept ? 1 : 0;
        if (st.acceptance_rate) st.acceptance_rate[chain] = alpha;
        if (st.log_density) st.log_density[chain] = lpn;
        if (st.hamiltonian_energy) st.hamiltonian_energy[chain] = H;
        if (st.hamiltonian_energy_error) st.hamiltonian_energy_error[chain] = H - H0;
        if (st.numerical_error) st.numerical_error[chain] = finite_d(H1) ? 0 : 1;
    }
}

#ifndef AHMC_SIMT_EMULATION  // host launch code (skipped by the CPU SIMT emulation harness, tests/simt_emu/)
// ---------------------------------------------------------------------------------------------
// dispatch
// ---------------------------------------------------------------------------------------------
#ifndef __CUDACC_RTC__  // host launch code (the kernels above are also compiled at run time for user targets, ahmc_user.cu)
// can the fast path of this launch use the lane-contiguous vector layout?  (full tile D == 32*E, rows and coefficient
// vectors aligned to the vector width)
template <int E>
static bool contig_ok(const LeapfrogArgs& a) {
    constexpr int V = E >= 4 ? 4 : E;
    const uintptr_t m = (uintptr_t)(8 * V - 1);
    auto al = [&](const void* p) { return ((uintptr_t)p & m) == 0; };
    if (a.D != 32 * E || a.ld_in % V || a.ld_out % V) return false;  // full tiles only: no bounds predicate in the kernel
    if (!al(a.th_in) || !al(a.r_in) || !al(a.g_in) || !al(a.th_out) || !al(a.r_out) || !al(a.g_out) || !al(a.dr_out)) return false;
    if (a.metric.kind == AHMC_METRIC_DIAG && (!al(a.metric.Minv) || a.metric.chain_stride % V)) return false;
    if (a.model.kind == AHMC_MODEL_DIAG_GAUSS && (!al(a.model.p0) || !al(a.model.p1))) return false;
    return true;
}

template <int MODEL, int METRIC, int G, int E, bool CONTIG>
static cudaError_t launch_lf_c(const LeapfrogArgs& a, cudaStream_t st) {
    const int chains_per_block = kBlockThreads / G;
    const long long blocks = (a.N + chains_per_block - 1) / chains_per_block;
    size_t sm = smem_bytes(MODEL, METRIC, a.D, G);
    const int occ = a.resident_blocks_per_sm;
    if (occ > 0) {
        // occupancy throttle (host-memory lanes): pad the dynamic shared memory so that only this many CTAs fit on
        // an SM; the grid then runs in staggered waves, some CTAs storing while others are still loading
        const size_t pad = (size_t)(227 * 1024) / (size_t)occ - 1024;
        if (pad > sm) sm = pad;
    }
    if (sm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(leapfrog_kernel<MODEL, METRIC, G, E, CONTIG>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
    }
    leapfrog_kernel<MODEL, METRIC, G, E, CONTIG><<<(unsigned)blocks, kBlockThreads, sm, st>>>(a);
    return cudaGetLastError();
}
template <int MODEL, int METRIC, int G, int E>
static cudaError_t launch_lf_t(const LeapfrogArgs& a, cudaStream_t st) {
    if constexpr (FastCapable<MODEL, METRIC>::value && G == 32 && E >= 2) {
        if (!(a.flags & AHMC_FLAG_EXACT_CHECKS) && !(a.temper_alpha > 0.0) && contig_ok<E>(a))
            return launch_lf_c<MODEL, METRIC, G, E, true>(a, st);
    }
    return launch_lf_c<MODEL, METRIC, G, E, false>(a, st);
}
template <int MODEL, int METRIC, int G, int E>
static cudaError_t launch_fe_t(const FindEpsArgs& a, cudaStream_t st) {
    const int chains_per_block = kBlockThreads / G;
    const long long blocks = (a.N + chains_per_block - 1) / chains_per_block;
    size_t sm = smem_bytes(MODEL, METRIC, a.D, G);
    if (sm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(find_eps_kernel<MODEL, METRIC, G, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
    }
    find_eps_kernel<MODEL, METRIC, G, E><<<(unsigned)blocks, kBlockThreads, sm, st>>>(a);
    return cudaGetLastError();
}
template <int MODEL, int METRIC, int G, int E>
static cudaError_t launch_pp_t(const PhasepointArgs& a, cudaStream_t st) {
    const int chains_per_block = kBlockThreads / G;
    const long long blocks = (a.N + chains_per_block - 1) / chains_per_block;
    size_t sm = smem_bytes(MODEL, METRIC, a.D, G);
    if (sm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(phasepoint_kernel<MODEL, METRIC, G, E>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
    }
    phasepoint_kernel<MODEL, METRIC, G, E><<<(unsigned)blocks, kBlockThreads, sm, st>>>(a);
    return cudaGetLastError();
}
template <int MODEL, int METRIC, int G, int E>
static cudaError_t launch_hmc_t(const HmcArgs& a, cudaStream_t st) {
    const int chains_per_block = kBlockThreads / G;
    const long long blocks = (a.lf.N + chains_per_block - 1) / chains_per_block;
    size_t sm = smem_bytes(MODEL, METRIC, a.lf.D, G);
    if (sm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(hmc_kernel<MODEL, METRIC, G, E>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
    }
    hmc_kernel<MODEL, METRIC, G, E><<<(unsigned)blocks, kBlockThreads, sm, st>>>(a);
    return cudaGetLastError();
}
template <int METRIC, int G, int E>
static cudaError_t launch_kd_t(const SplitArgs& a, cudaStream_t st) {
    const long long blocks = (a.N + kBlockThreads / G - 1) / (kBlockThreads / G);
    size_t sm = smem_bytes(AHMC_MODEL_STD_NORMAL, METRIC, a.D, G);
    if (sm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kick_drift_kernel<METRIC, G, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
    }
    kick_drift_kernel<METRIC, G, E><<<(unsigned)blocks, kBlockThreads, sm, st>>>(a);
    return cudaGetLastError();
}
template <int METRIC, int G, int E>
static cudaError_t launch_ke_t(const SplitArgs& a, cudaStream_t st) {
    const long long blocks = (a.N + kBlockThreads / G - 1) / (kBlockThreads / G);
    size_t sm = smem_bytes(AHMC_MODEL_STD_NORMAL, METRIC, a.D, G);
    if (sm > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kick_energy_kernel<METRIC, G, E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
        if (e != cudaSuccess) return e;
    }
    kick_energy_kernel<METRIC, G, E><<<(unsigned)blocks, kBlockThreads, sm, st>>>(a);
    return cudaGetLastError();
}
template <int DUMMY, int G, int E>
static cudaError_t launch_mh_t(const MhArgs& a, cudaStream_t st) {
    const long long blocks = (a.N + kBlockThreads / G - 1) / (kBlockThreads / G);
    mh_select_kernel<G, E><<<(unsigned)blocks, kBlockThreads, 0, st>>>(a);
    return cudaGetLastError();
}
template <int METRIC, int G, int E>
static cudaError_t launch_mom_t(const MomentumArgs& a, cudaStream_t st) {
    const int chains_per_block = kBlockThreads / G;
    const long long blocks = (a.N + chains_per_block - 1) / chains_per_block;
    momentum_kernel<METRIC, G, E><<<(unsigned)blocks, kBlockThreads, 0, st>>>(a);
    return cudaGetLastError();
}

#define AHMC_DISPATCH_LAYOUT(FN, ...)                                   \
    do {                                                                \
        if (G == 4 && E == 1) return FN<__VA_ARGS__, 4, 1>(a, st);      \
        if (G == 8 && E == 1) return FN<__VA_ARGS__, 8, 1>(a, st);      \
        if (G == 16 && E == 1) return FN<__VA_ARGS__, 16, 1>(a, st);    \
        if (G == 32 && E == 1) return FN<__VA_ARGS__, 32, 1>(a, st);    \
        if (G == 32 && E == 2) return FN<__VA_ARGS__, 32, 2>(a, st);    \
        if (G == 32 && E == 4) return FN<__VA_ARGS__, 32, 4>(a, st);    \
        if (G == 32 && E == 8) return FN<__VA_ARGS__, 32, 8>(a, st);    \
        if (G == 32 && E == 16) return FN<__VA_ARGS__, 32, 16>(a, st);  \
        return cudaErrorInvalidValue;                                   \
    } while (0)

template <int MODEL, int METRIC>
static cudaError_t lf_layout(const LeapfrogArgs& a, cudaStream_t st, int G, int E) {
    AHMC_DISPATCH_LAYOUT(launch_lf_t, MODEL, METRIC);
}
template <int MODEL, int METRIC>
static cudaError_t fe_layout(const FindEpsArgs& a, cudaStream_t st, int G, int E) {
    AHMC_DISPATCH_LAYOUT(launch_fe_t, MODEL, METRIC);
}
template <int MODEL, int METRIC>
static cudaError_t pp_layout(const PhasepointArgs& a, cudaStream_t st, int G, int E) {
    AHMC_DISPATCH_LAYOUT(launch_pp_t, MODEL, METRIC);
}
template <int MODEL, int METRIC>
static cudaError_t hmc_layout(const HmcArgs& a, cudaStream_t st, int G, int E) {
    AHMC_DISPATCH_LAYOUT(launch_hmc_t, MODEL, METRIC);
}
template <int METRIC>
static cudaError_t kd_layout(const SplitArgs& a, cudaStream_t st, int G, int E) {
    AHMC_DISPATCH_LAYOUT(launch_kd_t, METRIC);
}
template <int METRIC>
static cudaError_t ke_layout(const SplitArgs& a, cudaStream_t st, int G, int E) {
    AHMC_DISPATCH_LAYOUT(launch_ke_t, METRIC);
}
template <int DUMMY>
static cudaError_t mh_layout(const MhArgs& a, cudaStream_t st, int G, int E) {
    AHMC_DISPATCH_LAYOUT(launch_mh_t, DUMMY);
}
template <int METRIC>
static cudaError_t mom_layout(const MomentumArgs& a, cudaStream_t st, int G, int E) {
    AHMC_DISPATCH_LAYOUT(launch_mom_t, METRIC);
}

#define AHMC_DISPATCH_MM(FN, model_kind, metric_kind)                                                   \
    do {                                                                                                \
        switch ((model_kind) * 3 + (metric_kind)) {                                                     \
            case 0: return FN<AHMC_MODEL_STD_NORMAL, AHMC_METRIC_UNIT>(a, st, G, E);                    \
            case 1: return FN<AHMC_MODEL_STD_NORMAL, AHMC_METRIC_DIAG>(a, st, G, E);                    \
            case 2: return FN<AHMC_MODEL_STD_NORMAL, AHMC_METRIC_DENSE>(a, st, G, E);                   \
            case 3: return FN<AHMC_MODEL_DIAG_GAUSS, AHMC_METRIC_UNIT>(a, st, G, E);                    \
            case 4: return FN<AHMC_MODEL_DIAG_GAUSS, AHMC_METRIC_DIAG>(a, st, G, E);                    \
            case 5: return FN<AHMC_MODEL_DIAG_GAUSS, AHMC_METRIC_DENSE>(a, st, G, E);                   \
            case 6: return FN<AHMC_MODEL_DENSE_GAUSS, AHMC_METRIC_UNIT>(a, st, G, E);                   \
            case 7: return FN<AHMC_MODEL_DENSE_GAUSS, AHMC_METRIC_DIAG>(a, st, G, E);                   \
            case 8: return FN<AHMC_MODEL_DENSE_GAUSS, AHMC_METRIC_DENSE>(a, st, G, E);                  \
            case 9: return FN<AHMC_MODEL_FUNNEL, AHMC_METRIC_UNIT>(a, st, G, E);                        \
            case 10: return FN<AHMC_MODEL_FUNNEL, AHMC_METRIC_DIAG>(a, st, G, E);                       \
            case 11: return FN<AHMC_MODEL_FUNNEL, AHMC_METRIC_DENSE>(a, st, G, E);                      \
        }                                                                                               \
        return cudaErrorInvalidValue;                                                                   \
    } while (0)

cudaError_t launch_leapfrog(const LeapfrogArgs& a, cudaStream_t st, int* n_launches) {
    int G, E;
    if (a.D > 512) {  // beyond the register-resident layouts: the streaming form (ahmc_bigd.cu)
        if (n_launches) *n_launches += 1;
        return launch_leapfrog_big(a, st);
    }
    if (!pick_layout(a.D, &G, &E)) return cudaErrorInvalidValue;
    if (n_launches) *n_launches += 1;
    if (a.model.kind == AHMC_MODEL_USER) {  // run-time compiled kernels of a user target (ahmc_user.cu)
        const int cpb = kBlockThreads / G;
        return user_launch((UserModule*)a.model.user, UK_LEAPFROG, a.metric.kind, G, E, &a, (unsigned)((a.N + cpb - 1) / cpb),
                           smem_bytes(AHMC_MODEL_USER, a.metric.kind, a.D, G), st);
    }
    AHMC_DISPATCH_MM(lf_layout, a.model.kind, a.metric.kind);
}

cudaError_t launch_find_eps(const FindEpsArgs& a, cudaStream_t st, int* n_launches) {
    int G, E;
    if (!pick_layout(a.D, &G, &E)) return cudaErrorInvalidValue;
    if (n_launches) *n_launches += 1;
    if (a.model.kind == AHMC_MODEL_USER) {  // run-time compiled kernels of a user target (ahmc_user.cu)
        const int cpb = kBlockThreads / G;
        return user_launch((UserModule*)a.model.user, UK_FIND_EPS, a.metric.kind, G, E, &a, (unsigned)((a.N + cpb - 1) / cpb),
                           smem_bytes(AHMC_MODEL_USER, a.metric.kind, a.D, G), st);
    }
    AHMC_DISPATCH_MM(fe_layout, a.model.kind, a.metric.kind);
}

cudaError_t launch_phasepoint(const PhasepointArgs& a, cudaStream_t st, int* n_launches) {
    int G, E;
    if (a.D > 512) {
        if (n_launches) *n_launches += 1;
        return launch_phasepoint_big(a, st);
    }
    if (!pick_layout(a.D, &G, &E)) return cudaErrorInvalidValue;
    if (n_launches) *n_launches += 1;
    if (a.model.kind == AHMC_MODEL_USER) {  // run-time compiled kernels of a user target (ahmc_user.cu)
        const int cpb = kBlockThreads / G;
        return user_launch((UserModule*)a.model.user, UK_PHASEPOINT, a.metric.kind, G, E, &a, (unsigned)((a.N + cpb - 1) / cpb),
                           smem_bytes(AHMC_MODEL_USER, a.metric.kind, a.D, G), st);
    }
    AHMC_DISPATCH_MM(pp_layout, a.model.kind, a.metric.kind);
}

cudaError_t launch_hmc(const HmcArgs& a, cudaStream_t st, int* n_launches) {
    int G, E;
    if (!pick_layout(a.lf.D, &G, &E)) return cudaErrorInvalidValue;
    if (n_launches) *n_launches += 1;
    if (a.lf.model.kind == AHMC_MODEL_USER) {  // run-time compiled kernels of a user target (ahmc_user.cu)
        const int cpb = kBlockThreads / G;
        return user_launch((UserModule*)a.lf.model.user, UK_HMC, a.lf.metric.kind, G, E, &a, (unsigned)((a.lf.N + cpb - 1) / cpb),
                           smem_bytes(AHMC_MODEL_USER, a.lf.metric.kind, a.lf.D, G), st);
    }
    AHMC_DISPATCH_MM(hmc_layout, a.lf.model.kind, a.lf.metric.kind);
}

cudaError_t launch_kick_drift(const SplitArgs& a, cudaStream_t st, int* n_launches) {
    int G, E;
    if (!pick_layout(a.D, &G, &E)) return cudaErrorInvalidValue;
    if (n_launches) *n_launches += 1;
    switch (a.metric.kind) {
        case AHMC_METRIC_UNIT: return kd_layout<AHMC_METRIC_UNIT>(a, st, G, E);
        case AHMC_METRIC_DIAG: return kd_layout<AHMC_METRIC_DIAG>(a, st, G, E);
        case AHMC_METRIC_DENSE: return kd_layout<AHMC_METRIC_DENSE>(a, st, G, E);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_kick_energy(const SplitArgs& a, cudaStream_t st, int* n_launches) {
    int G, E;
    if (!pick_layout(a.D, &G, &E)) return cudaErrorInvalidValue;
    if (n_launches) *n_launches += 1;
    switch (a.metric.kind) {
        case AHMC_METRIC_UNIT: return ke_layout<AHMC_METRIC_UNIT>(a, st, G, E);
        case AHMC_METRIC_DIAG: return ke_layout<AHMC_METRIC_DIAG>(a, st, G, E);
        case AHMC_METRIC_DENSE: return ke_layout<AHMC_METRIC_DENSE>(a, st, G, E);
    }
    return cudaErrorInvalidValue;
}
cudaError_t launch_mh_select(const MhArgs& a, cudaStream_t st, int* n_launches) {
    int G, E;
    if (!pick_layout(a.D, &G, &E)) return cudaErrorInvalidValue;
    if (n_launches) *n_launches += 1;
    return mh_layout<0>(a, st, G, E);
}

cudaError_t launch_rand_momentum(const MomentumArgs& a, cudaStream_t st, int* n_launches) {
    int G, E;
    if (!pick_layout(a.D, &G, &E)) return cudaErrorInvalidValue;
    if (n_launches) *n_launches += 1;
    switch (a.metric.kind) {
        case AHMC_METRIC_UNIT: return mom_layout<AHMC_METRIC_UNIT>(a, st, G, E);
        case AHMC_METRIC_DIAG: return mom_layout<AHMC_METRIC_DIAG>(a, st, G, E);
        case AHMC_METRIC_DENSE: return mom_layout<AHMC_METRIC_DENSE>(a, st, G, E);
    }
    return cudaErrorInvalidValue;
}

#endif  // AHMC_SIMT_EMULATION

#endif  // __CUDACC_RTC__

}  // namespace ahmc
