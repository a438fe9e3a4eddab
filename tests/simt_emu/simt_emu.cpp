// simt_emu.cpp -- a minimal CPU emulator of CUDA's warp-synchronous execution model, enough to run the product's NUTS
// kernel SOURCE (advancedhmc.jl_b200/csrc/ahmc_nuts_kernel.cuh, unmodified) on the host: every CUDA thread of a block is
// a host thread; the 32 threads of a warp meet at a barrier in every shuffle / vote / __syncwarp, which is exactly the
// convergence the kernel's warp-uniform control flow guarantees on the device.  Blocks run one after another.
// TEST INFRASTRUCTURE ONLY (tests/test_simt_emulation.py): it lets kernel changes be checked against the oracle without
// a GPU; it says nothing about performance.
#include <pthread.h>

#include <atomic>
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

// run `kernel(args)` for a 1-D grid of `blocks` blocks of `threads` threads (threads % 32 == 0)
void emu_launch(void (*kernel)(const void*), const void* args, int blocks, int threads) {
    blockDim = {(unsigned)threads, 1, 1};
    gridDim = {(unsigned)blocks, 1, 1};
    const int nwarps = threads / 32;
    for (int b = 0; b < blocks; ++b) {
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
