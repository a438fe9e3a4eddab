// nuts_emu.cpp -- runs the product's NUTS kernel source (ahmc_nuts_kernel.cuh) under the CPU SIMT emulator.
// TEST INFRASTRUCTURE ONLY: built by tests/test_simt_emulation.py with g++; the kernel header is included unmodified
// (its host launch code is skipped with AHMC_SIMT_EMULATION).
#define AHMC_SIMT_EMULATION 1
#include <cstdlib>
#include <vector>

#include "ahmc_nuts_kernel.cuh"

namespace ahmc {
double smem[1 << 16];  // the block's dynamic shared memory (`extern __shared__ double smem[]` in the kernel)
}
void emu_launch(void (*kernel)(const void*), const void* args, int blocks, int threads);

using namespace ahmc;

struct EmuNuts {
    int32_t model_kind, metric_kind, D;
    int64_t N;
    const double *p0, *p1;
    double c0;
    const double* Minv;
    int64_t minv_stride;
    const double* cholU;
    double eps;
    const double* eps_chain;
    int32_t max_depth;
    double delta_max;
    int32_t sampler, criterion;
    uint64_t seed, offset;
    const double *normal_tape, *exp_tape;
    int64_t exp_stride;
    const uint8_t* dir_tape;
    int64_t dir_stride;
    double partial_alpha;
    int32_t refresh;
    const double *th_in, *r_in, *g_in, *lp_in;
    double *th_out, *r_out, *g_out, *lp_out, *lk_out;
    int32_t *n_steps, *tree_depth;
    uint8_t* numerical;
    double *acc, *dH, *dHmax;
    int32_t n_transitions;
    double* draws;
    int32_t adapt, n_adapts, init_buffer, term_buffer, window_size;
    double delta, gamma, t0, kappa;
    int32_t adapt_metric, n_min;
    double *eps_rw, *minv_rw, *eps_trace;
    double temper_alpha;
    int32_t coop_padded;  // 1: also hand the kernel column-padded copies of the dense matrices (one bulk copy per chunk)
};

template <int MODEL, int METRIC, int G, int E, bool VAR, bool ADAPT>
static constexpr bool coop_form() {  // like launch_nuts_v: dense operators, one chain per warp, default family
    return (MODEL == AHMC_MODEL_DENSE_GAUSS || METRIC == AHMC_METRIC_DENSE) && G == 32 && !VAR;
}
template <int MODEL, int METRIC, int G, int E, bool VAR, bool ADAPT>
static void thunk(const void* p) {
    const NutsArgs& a = *static_cast<const NutsArgs*>(p);
    if constexpr (coop_form<MODEL, METRIC, G, E, VAR, ADAPT>()) {
        if (E >= 2 && E <= 8 && a.D == G * E) nuts_kernel<MODEL, METRIC, G, E, VAR, ADAPT, true, true>(a);
        else nuts_kernel<MODEL, METRIC, G, E, VAR, ADAPT, false, true>(a);
        return;
    }
    if (G == 32 && E >= 2 && E <= 8 && a.D == G * E) {  // the full-tile instantiation (D a compile-time constant), like launch_nuts_v
        nuts_kernel<MODEL, METRIC, G, E, VAR, ADAPT, true>(a);
        return;
    }
    nuts_kernel<MODEL, METRIC, G, E, VAR, ADAPT, false>(a);
}

typedef void (*KernelFn)(const void*);

template <int MODEL, int METRIC, bool VAR, bool ADAPT>
static KernelFn by_layout(int G, int E) {
    if (G == 4 && E == 1) return thunk<MODEL, METRIC, 4, 1, VAR, ADAPT>;
    if (G == 8 && E == 1) return thunk<MODEL, METRIC, 8, 1, VAR, ADAPT>;
    if (G == 32 && E == 2) return thunk<MODEL, METRIC, 32, 2, VAR, ADAPT>;
    if (G == 16 && E == 1) return thunk<MODEL, METRIC, 16, 1, VAR, ADAPT>;
    if (G == 32 && E == 4) return thunk<MODEL, METRIC, 32, 4, VAR, ADAPT>;
    if (G == 32 && E == 8) return thunk<MODEL, METRIC, 32, 8, VAR, ADAPT>;
    if constexpr (MODEL == AHMC_MODEL_DENSE_GAUSS && METRIC == AHMC_METRIC_DENSE && !VAR && !ADAPT) {
        if (G == 32 && E == 16) return thunk<MODEL, METRIC, 32, 16, VAR, ADAPT>;  // D in (256, 512]: 16-column chunks, 2 stages
    }
    return nullptr;
}

template <bool VAR, bool ADAPT>
static KernelFn by_model(int model, int metric, int G, int E) {
    if (model == AHMC_MODEL_STD_NORMAL && metric == AHMC_METRIC_UNIT) return by_layout<AHMC_MODEL_STD_NORMAL, AHMC_METRIC_UNIT, VAR, ADAPT>(G, E);
    if (model == AHMC_MODEL_DIAG_GAUSS && metric == AHMC_METRIC_DIAG) return by_layout<AHMC_MODEL_DIAG_GAUSS, AHMC_METRIC_DIAG, VAR, ADAPT>(G, E);
    if (model == AHMC_MODEL_FUNNEL && metric == AHMC_METRIC_DIAG) return by_layout<AHMC_MODEL_FUNNEL, AHMC_METRIC_DIAG, VAR, ADAPT>(G, E);
    if (ADAPT && !VAR && model == AHMC_MODEL_DENSE_GAUSS && metric == AHMC_METRIC_DIAG)  // adaptive family, cooperative form
        return by_layout<AHMC_MODEL_DENSE_GAUSS, AHMC_METRIC_DIAG, false, true>(G, E);
    if (!ADAPT) {
        if (model == AHMC_MODEL_DENSE_GAUSS && metric == AHMC_METRIC_DENSE) return by_layout<AHMC_MODEL_DENSE_GAUSS, AHMC_METRIC_DENSE, VAR, false>(G, E);
        if (model == AHMC_MODEL_DIAG_GAUSS && metric == AHMC_METRIC_UNIT) return by_layout<AHMC_MODEL_DIAG_GAUSS, AHMC_METRIC_UNIT, VAR, false>(G, E);
        if (!VAR) {  // the other two dense-operator combinations of the COOP form
            if (model == AHMC_MODEL_DENSE_GAUSS && metric == AHMC_METRIC_DIAG) return by_layout<AHMC_MODEL_DENSE_GAUSS, AHMC_METRIC_DIAG, false, false>(G, E);
            if (model == AHMC_MODEL_DIAG_GAUSS && metric == AHMC_METRIC_DENSE) return by_layout<AHMC_MODEL_DIAG_GAUSS, AHMC_METRIC_DENSE, false, false>(G, E);
        }
    }
    return nullptr;
}

// the product's host-side window schedule (ahmc_kernels.cuh), for the comparison with the oracle's
extern "C" int emu_window_schedule(int init_buffer, int term_buffer, int window_size, int n_adapts, int* ws, int* we, int* splits) {
    AdaptDev ad{};
    if (!stan_window_schedule(ad, init_buffer, term_buffer, window_size, n_adapts)) return -1;
    *ws = ad.window_start;
    *we = ad.window_end;
    for (int i = 0; i < ad.n_splits; ++i) splits[i] = ad.splits[i];
    return ad.n_splits;
}

extern "C" int emu_nuts(const EmuNuts* q) {
    int G, E;
    const int D = q->D;
    if (D <= 4) G = 4, E = 1;
    else if (D <= 8) G = 8, E = 1;
    else if (D <= 16) G = 16, E = 1;
    else if (D > 32 && D <= 64) G = 32, E = 2;
    else if (D > 64 && D <= 128) G = 32, E = 4;  // the headline layout
    else if (D > 128 && D <= 256) G = 32, E = 8;  // C5's layout
    else if (D > 256 && D <= 512) G = 32, E = 16;
    else return -1;
    NutsArgs a{};
    a.model = ModelDev{q->model_kind, D, q->p0, q->p1, q->c0};
    a.metric = MetricDev{q->metric_kind, q->Minv, q->minv_stride, q->cholU};
    std::vector<double> pad_p, pad_m, pad_u;
    if (q->coop_padded) {  // what ahmc_model_create / nuts_impl prepare on the device (launch_pad_columns)
        const int lds = coop_lds(D);
        auto pad = [&](const double* A, std::vector<double>& out) {
            out.assign((size_t)D * lds, 0.0);
            for (int k = 0; k < D; ++k)
                for (int r = 0; r < D; ++r) out[(size_t)k * lds + r] = A[(size_t)k * D + r];
            return out.data();
        };
        if (q->model_kind == AHMC_MODEL_DENSE_GAUSS) a.model.p1_coop = pad(q->p1, pad_p);
        if (q->metric_kind == AHMC_METRIC_DENSE) {
            a.metric.Minv_coop = pad(q->Minv, pad_m);
            if (q->cholU) a.metric.cholU_coop = pad(q->cholU, pad_u);
        }
    }
    a.D = D;
    a.N = q->N;
    a.eps = q->eps;
    a.eps_chain = q->eps_chain;
    a.max_depth = q->max_depth;
    a.delta_max = q->delta_max;
    a.sampler = q->sampler;
    a.criterion = q->criterion;
    a.rng = RngDev{q->seed, q->offset, q->normal_tape, q->exp_tape, q->exp_stride, q->dir_tape, q->dir_stride, q->partial_alpha, q->temper_alpha};
    a.refresh = q->refresh;
    a.th_in = q->th_in; a.r_in = q->r_in; a.g_in = q->g_in; a.lp_in = q->lp_in;
    a.ld_in = D;
    a.th_out = q->th_out; a.r_out = q->r_out; a.g_out = q->g_out; a.lp_out = q->lp_out; a.lk_out = q->lk_out;
    a.ld_out = D;
    a.st = StatsDev{};
    a.st.n_steps = q->n_steps;
    a.st.tree_depth = q->tree_depth;
    a.st.numerical_error = q->numerical;
    a.st.acceptance_rate = q->acc;
    a.st.hamiltonian_energy_error = q->dH;
    a.st.max_hamiltonian_energy_error = q->dHmax;
    a.n_transitions = q->n_transitions;
    a.draws = q->draws;
    if (q->adapt) {
        AdaptDev& ad = a.ad;
        ad.enabled = 1;
        ad.n_adapts = q->n_adapts;
        ad.delta = q->delta; ad.gamma = q->gamma; ad.t0 = q->t0; ad.kappa = q->kappa;
        ad.adapt_metric = q->adapt_metric;
        ad.n_min = q->n_min;
        if (!stan_window_schedule(ad, q->init_buffer, q->term_buffer, q->window_size, q->n_adapts)) return -3;  // as nuts_impl
        ad.eps = q->eps_rw;
        a.eps_chain = q->eps_rw;
        ad.minv = q->minv_rw;
        ad.eps_trace = q->eps_trace;
    }
    const long long stride = nuts_level_doubles(D, q->max_depth) + (q->adapt ? 2LL * D : 0);
    std::vector<double> scratch((size_t)stride * (size_t)q->N, 0.0);
    a.scratch = scratch.data();
    a.scratch_stride = stride;
    const bool var = q->sampler != 0 || q->criterion != 0;
    KernelFn fn = q->adapt ? by_model<false, true>(q->model_kind, q->metric_kind, G, E)
                           : (var ? by_model<true, false>(q->model_kind, q->metric_kind, G, E)
                                  : by_model<false, false>(q->model_kind, q->metric_kind, G, E));
    if (!fn) return -2;
    const bool dense_ops = q->model_kind == AHMC_MODEL_DENSE_GAUSS || q->metric_kind == AHMC_METRIC_DENSE;
    const bool coop = dense_ops && G == 32 && !var;  // blocks of kCoopWarps chains sharing the D x D products
    const int threads = coop ? kCoopThreads : kBlockThreads;
    const int chains_per_block = threads / G;
    const int blocks = (int)((q->N + chains_per_block - 1) / chains_per_block);
    emu_launch(fn, &a, blocks, threads);
    return 0;
}
