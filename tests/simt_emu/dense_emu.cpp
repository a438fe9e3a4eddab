// dense_emu.cpp -- K4, the tile-per-CTA trajectory kernel for GEMM-shaped operators (advancedhmc.jl_b200/csrc/
// ahmc_dense.cu: `dense_traj_kernel`, `pad_norm_kernel`, `vec_norm_kernel`, unmodified) under the CPU SIMT emulator.
// The kernel's five PTX wrappers are restated here with the same contracts:
//   * an mbarrier is (completed phases, pending arrivals, pending transaction bytes); `mbar_wait(parity)` returns once
//     the phase of that parity has completed;
//   * `bulk_g2s` copies synchronously and completes its bytes on the barrier;
//   * `dmma` is `mma.sync.aligned.m8n8k4.row.col.f64`: lane l holds A[l/4][l%4], B[l%4][l/4] and C[l/4][2(l%4)+{0,1}].
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
// An mbarrier lives in the kernel's shared memory as one 64-bit word; here: bits 0..19 completed phases, 20..31 pending
// arrivals of the current phase, 32..63 pending transaction bytes (biased by 2^31: complete_tx may precede expect_tx).
// The expected arrival count of each barrier is kept beside it (indexed by its address: at most 8 barriers per block).
namespace {
constexpr uint64_t kTxBias = 1ull << 31;
struct BarInfo {
    uint64_t* bar;
    uint32_t count;
};
BarInfo bar_info[8];
std::atomic<int> n_bar_info{0};
uint32_t expected_arrivals(uint64_t* bar) {
    for (int i = 0; i < n_bar_info.load(); ++i)
        if (bar_info[i].bar == bar) return bar_info[i].count;
    return 0;  // unreachable for an initialised barrier
}
// apply (arrivals, +/- bytes) atomically; complete the phase when both reach zero
void bar_update(uint64_t* bar, uint32_t arrivals, int64_t tx) {
    std::atomic_ref<uint64_t> b(*bar);
    uint64_t old = b.load(), neu;
    do {
        uint64_t phases = old & 0xfffffu, pend = (old >> 20) & 0xfffu;
        int64_t bytes = (int64_t)(old >> 32) - (int64_t)kTxBias + tx;
        pend -= arrivals;
        if (pend == 0 && bytes == 0) {
            phases = (phases + 1) & 0xfffffu;
            pend = expected_arrivals(bar);
        }
        neu = phases | (pend << 20) | ((uint64_t)(bytes + (int64_t)kTxBias) << 32);
    } while (!b.compare_exchange_weak(old, neu));
}
}  // namespace
void mbar_init(uint64_t* bar, int count) {  // thread 0 only, before the block barrier
    int i = 0;
    for (; i < n_bar_info.load(); ++i)
        if (bar_info[i].bar == bar) break;
    if (i == n_bar_info.load()) n_bar_info.store(i + 1);
    bar_info[i] = BarInfo{bar, (uint32_t)count};
    std::atomic_ref<uint64_t>(*bar).store(((uint64_t)count << 20) | (kTxBias << 32));
}
void mbar_fence_init() {}
void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { bar_update(bar, 1, (int64_t)bytes); }
void mbar_arrive(uint64_t* bar) { bar_update(bar, 1, 0); }
void mbar_wait(uint64_t* bar, uint32_t parity) {
    while ((uint32_t)(std::atomic_ref<uint64_t>(*bar).load() & 1u) == parity) std::this_thread::yield();
}
void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    std::memcpy(dst, src, bytes);
    bar_update(bar, 0, -(int64_t)bytes);
}
void dmma(double& d0, double& d1, double a, double b) {
    double A[32], B[32];
    emu_gather2(a, b, A, B);
    const int lane = emu_lane(), row = lane >> 2, c0 = 2 * (lane & 3);
    for (int k = 0; k < 4; ++k) {
        d0 = fma(A[row * 4 + k], B[c0 * 4 + k], d0);
        d1 = fma(A[row * 4 + k], B[(c0 + 1) * 4 + k], d1);
    }
}
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
    n_bar_info.store(0);
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
