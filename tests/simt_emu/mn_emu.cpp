// mn_emu.cpp -- the MultinomialTS static-transition kernel (advancedhmc.jl_b200/csrc/ahmc_multinomial.cu, unmodified)
// under the CPU SIMT emulator.  TEST INFRASTRUCTURE ONLY (tests/test_simt_emulation.py).
#define AHMC_SIMT_EMULATION 1
#include <vector>

#include "ahmc_multinomial.cu"

namespace ahmc {
double smem[1 << 16];
}
void emu_launch(void (*kernel)(const void*), const void* args, int blocks, int threads);
using namespace ahmc;

struct EmuMn {
    int32_t model_kind, metric_kind, D;
    int64_t N;
    const double *p0, *p1;
    const double* Minv;
    const double* cholU;
    double eps;
    int32_t n_steps, n_fwd;
    const double *normal_tape, *unif_tape;
    const double *th_in, *g_in, *lp_in;
    double *th_out, *r_out, *g_out, *lp_out, *lk_out;
    double* acc;
    int32_t* index;
    double temper_alpha;
};

template <int MODEL, int METRIC, int G, int E>
static void mn_thunk(const void* p) { multinomial_kernel<MODEL, METRIC, G, E>(*static_cast<const MultinomialArgs*>(p)); }
typedef void (*KernelFn)(const void*);
template <int MODEL, int METRIC>
static KernelFn pick(int G, int E) {
    if (G == 4 && E == 1) return mn_thunk<MODEL, METRIC, 4, 1>;
    if (G == 8 && E == 1) return mn_thunk<MODEL, METRIC, 8, 1>;
    if (G == 32 && E == 2) return mn_thunk<MODEL, METRIC, 32, 2>;
    return nullptr;
}

extern "C" int emu_multinomial(const EmuMn* q) {
    int G, E;
    const int D = q->D;
    if (D <= 4) G = 4, E = 1;
    else if (D <= 8) G = 8, E = 1;
    else if (D > 32 && D <= 64) G = 32, E = 2;
    else return -1;
    std::vector<double> r_in((size_t)D * q->N, 0.0), energies((size_t)(q->n_steps + 1) * q->N, 0.0);
    MultinomialArgs a{};
    a.model = ModelDev{q->model_kind, D, q->p0, q->p1, 0.0};
    a.metric = MetricDev{q->metric_kind, q->Minv, 0, q->cholU};
    a.D = D;
    a.N = q->N;
    a.eps = q->eps;
    a.n_steps = q->n_steps;
    a.n_fwd = q->n_fwd;
    a.refresh = 1;
    a.rng = RngDev{1, 0, q->normal_tape, q->unif_tape, 1, nullptr, 0, 0.0, q->temper_alpha};
    a.th_in = q->th_in; a.r_in = r_in.data(); a.g_in = q->g_in; a.lp_in = q->lp_in;
    a.ld_in = D;
    a.th_out = q->th_out; a.r_out = q->r_out; a.g_out = q->g_out; a.lp_out = q->lp_out; a.lk_out = q->lk_out;
    a.ld_out = D;
    a.st = StatsDev{};
    a.st.acceptance_rate = q->acc;
    a.st.tree_depth = q->index;
    a.energies = energies.data();
    KernelFn fn = nullptr;
    const int m = q->model_kind, me = q->metric_kind;
    if (m == AHMC_MODEL_DIAG_GAUSS && me == AHMC_METRIC_DIAG) fn = pick<AHMC_MODEL_DIAG_GAUSS, AHMC_METRIC_DIAG>(G, E);
    else if (m == AHMC_MODEL_FUNNEL && me == AHMC_METRIC_DENSE) fn = pick<AHMC_MODEL_FUNNEL, AHMC_METRIC_DENSE>(G, E);
    else if (m == AHMC_MODEL_DENSE_GAUSS && me == AHMC_METRIC_UNIT) fn = pick<AHMC_MODEL_DENSE_GAUSS, AHMC_METRIC_UNIT>(G, E);
    if (!fn) return -2;
    const int chains_per_block = kBlockThreads / G;
    emu_launch(fn, &a, (int)((q->N + chains_per_block - 1) / chains_per_block), kBlockThreads);
    return 0;
}
