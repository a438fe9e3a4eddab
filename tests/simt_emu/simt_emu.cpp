// simt_emu.cpp -- a minimal CPU emulator of CUDA's warp-synchronous execution model, enough to run the product's NUTS
// kernel SOURCE (advancedhmc.jl_b200/csrc/ahmc_nuts_kernel.cuh, unmodified) on the host: every CUDA thread of a block is
// a host thread; the 32 threads of a warp meet at a barrier in every shuffle / vote / __syncwarp, which is exactly the
// convergence the kernel's warp-uniform control flow guarantees on the device.  Blocks run one after another.
// TEST INFRASTRUCTURE ONLY (tests/test_simt_emulation.py): it lets kernel changes be checked against the oracle without
// a GPU; it says nothing about performance.
#include <pthread.h>

#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "cuda_runtime.h"

thread_local EmuUint3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0};
EmuUint3 blockDim = {1, 1, 1}, gridDim = {1, 1, 1};

namespace {
struct Warp {
    pthread_barrier_t bar;
    uint64_t slot[32];
    double slot_b[32];
    unsigned pred[32];
};
thread_local Warp* tls_warp = nullptr;
thread_local int tls_lane = 0;
thread_local pthread_barrier_t* tls_block_bar = nullptr;
}  // namespace

void __syncthreads() { pthread_barrier_wait(tls_block_bar); }
static std::atomic<int> g_block_or{0};  // blocks run one at a time
int __syncthreads_or(int pred) {
    if (pred) g_block_or.store(1);
    pthread_barrier_wait(tls_block_bar);
    const int r = g_block_or.load();
    pthread_barrier_wait(tls_block_bar);
    if (threadIdx.x == 0) g_block_or.store(0);
    pthread_barrier_wait(tls_block_bar);
    return r;
}
unsigned atomicAdd(unsigned* addr, unsigned v) { return std::atomic_ref<unsigned>(*addr).fetch_add(v); }

int emu_lane() { return tls_lane; }
void emu_syncwarp() { pthread_barrier_wait(&tls_warp->bar); }
unsigned emu_ballot(bool p) {
    Warp* w = tls_warp;
    w->pred[tls_lane] = p ? 1u : 0u;
    pthread_barrier_wait(&w->bar);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= w->pred[i] << i;
    pthread_barrier_wait(&w->bar);
    return m;
}
uint64_t emu_shfl(uint64_t bits, int src_lane) {
    Warp* w = tls_warp;
    w->slot[tls_lane] = bits;
    pthread_barrier_wait(&w->bar);
    const uint64_t out = w->slot[src_lane & 31];
    pthread_barrier_wait(&w->bar);
    return out;
}
unsigned emu_reduce_max(unsigned mask, unsigned v) {
    Warp* w = tls_warp;
    w->slot[tls_lane] = v;
    pthread_barrier_wait(&w->bar);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i)
        if ((mask >> i) & 1u) m = w->slot[i] > m ? (unsigned)w->slot[i] : m;
    pthread_barrier_wait(&w->bar);
    return m;
}
int atomicOr(int* addr, int v) { return std::atomic_ref<int>(*addr).fetch_or(v); }
void emu_gather2(double a, double b, double* a32, double* b32) {
    Warp* w = tls_warp;
    std::memcpy(&w->slot[tls_lane], &a, 8);
    w->slot_b[tls_lane] = b;
    pthread_barrier_wait(&w->bar);
    std::memcpy(a32, w->slot, 32 * 8);
    std::memcpy(b32, w->slot_b, 32 * 8);
    pthread_barrier_wait(&w->bar);
}
int atomicMin(int* addr, int v) {
    std::atomic_ref<int> a(*addr);
    int old = a.load();
    while (old > v && !a.compare_exchange_weak(old, v)) {
    }
    return old;
}

namespace ahmc {
void emu_reset_barriers();  // (below) forget the previous block's mbarriers
}
// run `kernel(args)` for a 1-D grid of `blocks` blocks of `threads` threads (threads % 32 == 0)
void emu_launch(void (*kernel)(const void*), const void* args, int blocks, int threads) {
    blockDim = {(unsigned)threads, 1, 1};
    gridDim = {(unsigned)blocks, 1, 1};
    const int nwarps = threads / 32;
    for (int b = 0; b < blocks; ++b) {
        ahmc::emu_reset_barriers();
        std::vector<Warp> warps(nwarps);
        for (auto& w : warps) pthread_barrier_init(&w.bar, nullptr, 32);
        pthread_barrier_t block_bar;
        pthread_barrier_init(&block_bar, nullptr, threads);
        std::vector<std::thread> ts;
        ts.reserve(threads);
        for (int t = 0; t < threads; ++t)
            ts.emplace_back([&, t, b] {
                threadIdx = {(unsigned)t, 0, 0};
                blockIdx = {(unsigned)b, 0, 0};
                tls_warp = &warps[t / 32];
                tls_lane = t % 32;
                tls_block_bar = &block_bar;
                kernel(args);
            });
        for (auto& th : ts) th.join();
        for (auto& w : warps) pthread_barrier_destroy(&w.bar);
        pthread_barrier_destroy(&block_bar);
    }
}

// ---- mbarrier / bulk copy / fp64 MMA contracts of the kernels' PTX wrappers (ahmc_device.cuh) --------------------------
//   * an mbarrier is (completed phases, pending arrivals, pending transaction bytes); `mbar_wait(parity)` returns once
//     the phase of that parity has completed; `mbar_init` on an invalidated barrier re-arms it;
//   * `bulk_g2s` copies synchronously and completes its bytes on the barrier;
//   * `dmma` is `mma.sync.aligned.m8n8k4.row.col.f64`: lane l holds A[l/4][l%4], B[l%4][l/4] and C[l/4][2(l%4)+{0,1}].
namespace ahmc {
// An mbarrier lives in the kernel's shared memory as one 64-bit word; here: bits 0..19 completed phases, 20..31 pending
// arrivals of the current phase, 32..63 pending transaction bytes (biased by 2^31: complete_tx may precede expect_tx).
// The expected arrival count of each barrier is kept beside it (indexed by its address: at most 8 barriers per block).
namespace {
constexpr uint64_t kTxBias = 1ull << 31;
struct BarInfo {
    uint64_t* bar;
    uint32_t count;
};
BarInfo bar_info[16];
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
void emu_reset_barriers() { n_bar_info.store(0); }
void mbar_init(uint64_t* bar, int count) {  // thread 0 only, before the block barrier
    int i = 0;
    for (; i < n_bar_info.load(); ++i)
        if (bar_info[i].bar == bar) break;
    if (i == n_bar_info.load()) {
        if (i == 16) {
            std::fprintf(stderr, "[simt_emu] more than 16 mbarriers in one block\n");
            std::abort();
        }
        n_bar_info.store(i + 1);
    }
    bar_info[i] = BarInfo{bar, (uint32_t)count};
    std::atomic_ref<uint64_t>(*bar).store(((uint64_t)count << 20) | (kTxBias << 32));
}
void mbar_fence_init() {}
void mbar_inval(uint64_t*) {}
void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { bar_update(bar, 1, (int64_t)bytes); }
void mbar_arrive(uint64_t* bar) { bar_update(bar, 1, 0); }
void mbar_wait(uint64_t* bar, uint32_t parity) {
    // spin briefly, then sleep: a block is 256 host threads on a handful of cores, and waiters must not starve the workers
    for (int spins = 0; (uint32_t)(std::atomic_ref<uint64_t>(*bar).load() & 1u) == parity; ++spins) {
        if (spins < 16) std::this_thread::yield();
        else std::this_thread::sleep_for(std::chrono::microseconds(50));
        if (spins == 400000) {  // ~20 s: a protocol error in the kernel under test -- say where instead of hanging the suite
            const uint64_t v = std::atomic_ref<uint64_t>(*bar).load();
            std::fprintf(stderr, "[simt_emu] mbar_wait stuck: thread %u block %u barrier %p parity %u state phases=%llu pending=%llu tx=%lld\n",
                         threadIdx.x, blockIdx.x, (void*)bar, parity, (unsigned long long)(v & 0xfffffu),
                         (unsigned long long)((v >> 20) & 0xfffu), (long long)(v >> 32) - (long long)kTxBias);
            std::abort();
        }
    }
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
