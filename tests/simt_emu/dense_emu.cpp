// dense_emu.cpp -- K4, the tile-per-CTA trajectory kernel for GEMM-shaped operators (advancedhmc.jl_b200/csrc/
// ahmc_dense.cu: `dense_traj_kernel`, `pad_norm_kernel`, `vec_norm_kernel`, unmodified) under the CPU SIMT emulator.
// The kernel's PTX wrappers (mbarrier, bulk copy, DMMA) are restated in simt_emu.cpp with the same contracts.
// TEST INFRASTRUCTURE ONLY (tests/test_simt_emulation.py).
#define AHMC_SIMT_EMULATION 1
#define __shared__ static  // static shared variables; blocks run one at a time
#include <atomic>
#include <thread>
#include <vector>

#include "ahmc_dense.cu"

void emu_launch(void (*kernel)(const void*), const void* args, int blocks, int threads);

namespace ahmc {
unsigned char* emu_dynamic_smem = nullptr;
// (the mbarrier / bulk-copy / DMMA contracts live in simt_emu.cpp: the cooperative NUTS products use them too)
}  // namespace ahmc
using namespace ahmc;

template <int RB, int CB, int MINB>
static void dense_thunk(const void* p) { dense_traj_kernel<RB, CB, MINB>(*static_cast<const DenseArgs*>(p)); }
struct PadArgs {
    const double* A;
    int D, Dp;
    double* Ap;
    double* norm;
};
static void pad_thunk(const void* p) {
    const PadArgs& a = *static_cast<const PadArgs*>(p);
    pad_norm_kernel(a.A, a.D, a.Dp, a.Ap, a.norm);
}
static void vec_thunk(const void* p) {
    const PadArgs& a = *static_cast<const PadArgs*>(p);
    vec_norm_kernel(a.A, a.D, a.norm);
}

struct EmuDense {
    int32_t D;
    int64_t N;
    const double *P, *w, *mu;  // P: D x D column-major precision or null; w: 1/s^2 or null
    double c0;
    const double *Minv, *Mdiag;  // Minv: D x D column-major or null
    double eps;
    const double* eps_chain;
    int32_t n_steps, fwd;
    const double *th_in, *r_in, *g_in;
    double *th_out, *r_out, *g_out, *dr_out, *lp_out, *lk_out;
    uint32_t* status;
    int32_t* steps_done;
    uint8_t* need_exact;
    int32_t wide_tile;  // RB == 2 only: 1 -> the <2,4> (32-chain) form instead of the default <2,2,2>
    double norms_out[2];
};

extern "C" int emu_dense_stages() { return kStages; }
extern "C" int emu_dense(EmuDense* q) {
    // (mbarrier table: reset per block by emu_launch)
    int Dp, RB, CB;
    if (!dense_tile_shape(q->D, &Dp, &RB, &CB)) return -1;
    std::vector<double> Pp, Mp;
    double norms[2] = {0.0, 0.0};
    PadArgs pa{};
    if (q->Minv) {
        Mp.assign(dense_mat_doubles(Dp), -1.0);
        pa = PadArgs{q->Minv, q->D, Dp, Mp.data(), &norms[0]};
        emu_launch(pad_thunk, &pa, 1, 256);
    } else {
        pa = PadArgs{q->Mdiag, q->D, Dp, nullptr, &norms[0]};
        emu_launch(vec_thunk, &pa, 1, 32);
    }
    if (q->P) {
        Pp.assign(dense_mat_doubles(Dp), -1.0);
        pa = PadArgs{q->P, q->D, Dp, Pp.data(), &norms[1]};
        emu_launch(pad_thunk, &pa, 1, 256);
    } else {
        pa = PadArgs{q->w, q->D, Dp, nullptr, &norms[1]};
        emu_launch(vec_thunk, &pa, 1, 32);
    }
    q->norms_out[0] = norms[0];
    q->norms_out[1] = norms[1];
    DenseArgs a{};
    a.D = q->D; a.Dp = Dp; a.N = q->N; a.P = q->P ? Pp.data() : nullptr; a.w = q->w; a.mu = q->mu; a.c0 = q->c0;
    a.Minv = q->Minv ? Mp.data() : nullptr; a.Mdiag = q->Mdiag; a.norms = norms; a.eps = q->eps; a.eps_chain = q->eps_chain;
    a.n_steps = q->n_steps; a.fwd = q->fwd; a.th_in = q->th_in; a.r_in = q->r_in; a.g_in = q->g_in; a.ld_in = q->D;
    a.th_out = q->th_out; a.r_out = q->r_out; a.g_out = q->g_out; a.dr_out = q->dr_out; a.lp_out = q->lp_out; a.lk_out = q->lk_out;
    a.ld_out = q->D; a.status = q->status; a.steps_done = q->steps_done; a.need_exact = q->need_exact;
    void (*fn)(const void*) = nullptr;
    int CT = 0;
    if (RB == 1) fn = dense_thunk<1, 4, 1>, CT = 32;
    else if (RB == 2 && q->wide_tile) fn = dense_thunk<2, 4, 1>, CT = 32;
    else if (RB == 2) fn = dense_thunk<2, 2, 2>, CT = 16;
    else if (RB == 3) fn = dense_thunk<3, 2, 1>, CT = 16;
    else return -2;
    const int Ds = Dp + 4;
    const size_t sm = ((size_t)kStages * kKC * Ds + (size_t)CT * Ds + 8 * CT * 2) * sizeof(double) + 64;  // as launch_dense_t
    std::vector<double> smem(sm / sizeof(double) + 2, 0.0);
    emu_dynamic_smem = reinterpret_cast<unsigned char*>(smem.data());
    emu_launch(fn, &a, (int)((q->N + CT - 1) / CT), kDenseThreads);
    emu_dynamic_smem = nullptr;
    return 0;
}
