// Fake <cuda_runtime.h> for the CPU SIMT emulation harness (tests/simt_emu/): just enough host-side declarations for the
// product's kernel headers to compile with g++.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#ifndef __shared__
#define __shared__  // kernels that only use `extern __shared__` arrays; a harness for kernels with STATIC shared
#endif              // variables defines it as `static` before including this header (blocks run one at a time)
#define __launch_bounds__(...)

typedef int cudaError_t;
typedef void* cudaStream_t;
constexpr cudaError_t cudaSuccess = 0, cudaErrorInvalidValue = 1;

struct EmuUint3 {
    unsigned x, y, z;
};
extern thread_local EmuUint3 threadIdx, blockIdx;
extern EmuUint3 blockDim, gridDim;

// ---- warp-synchronous primitives: 32 host threads per warp in lock-step (tests/simt_emu/simt_emu.cpp)
unsigned emu_ballot(bool pred);
uint64_t emu_shfl(uint64_t bits, int src_lane);
void emu_syncwarp();
inline unsigned __ballot_sync(unsigned, bool pred) { return emu_ballot(pred); }
inline bool __any_sync(unsigned, bool pred) { return emu_ballot(pred) != 0u; }
inline bool __all_sync(unsigned, bool pred) { return emu_ballot(pred) == 0xffffffffu; }
inline void __syncwarp(unsigned = 0xffffffffu) { emu_syncwarp(); }
int emu_lane();
template <class T>
inline T emu_shfl_t(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    uint64_t b = 0;
    std::memcpy(&b, &v, sizeof(T));
    b = emu_shfl(b, src);
    T out;
    std::memcpy(&out, &b, sizeof(T));
    return out;
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lanemask, int = 32) { return emu_shfl_t(v, emu_lane() ^ lanemask); }
template <class T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    const int lane = emu_lane();
    return emu_shfl_t(v, (lane & ~(width - 1)) | (src & (width - 1)));
}

// ---- scalar intrinsics
template <class T>
inline T __ldg(const T* p) { return *p; }
inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32); }
inline int __double2hiint(double x) {
    uint64_t b;
    std::memcpy(&b, &x, 8);
    return (int)(b >> 32);
}
inline double __hiloint2double(int hi, int lo) {
    uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x;
    std::memcpy(&x, &b, 8);
    return x;
}
struct double2 {
    double x, y;
};
inline double2 make_double2(double x, double y) { return double2{x, y}; }
int atomicMin(int* addr, int v);
int atomicOr(int* addr, int v);
// every lane contributes (a, b); every lane receives all 32 of each (one exchange instead of 12 shuffles: emulated DMMA)
void emu_gather2(double a, double b, double* a32, double* b32);
using std::fabs;
using std::fma;
using std::fmax;
using std::fmin;
inline void sincospi(double x, double* s, double* c) {  // CUDA math API; only the Philox (non-tape) momentum draw uses it
    *s = std::sin(M_PI * x);
    *c = std::cos(M_PI * x);
}
// warp reduction over the lanes named in `mask` (sm_80+ __reduce_max_sync): every lane of the warp participates here
unsigned emu_reduce_max(unsigned mask, unsigned v);
inline int __reduce_max_sync(unsigned mask, int v) { return (int)(emu_reduce_max(mask, (unsigned)v ^ 0x80000000u) ^ 0x80000000u); }
inline unsigned __reduce_max_sync(unsigned mask, unsigned v) { return emu_reduce_max(mask, v); }

// ---- block-level primitives
void __syncthreads();
int __syncthreads_or(int pred);
inline void __threadfence() {}
unsigned atomicAdd(unsigned* addr, unsigned v);
