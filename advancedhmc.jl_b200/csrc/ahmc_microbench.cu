// ahmc_microbench.cu -- measurement helpers for bench.py (NOT part of the drop-in boundary; built into its own
// libahmc_microbench.so).  SURVEY 8(d): the fused trajectory kernel is bounded by the fp64 FMA pipe, whose measured
// peak is not in MEASURED_PEAKS.json, so it is measured here: every thread runs 8 independent DFMA chains.
#include <cuda_runtime.h>

#include <cstdio>

namespace {

template <int CHAINS>
__global__ void __launch_bounds__(256) dfma_kernel(double* sink, int iters, double a, double b) {
    double x[CHAINS];
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) x[j] = 1e-3 * (threadIdx.x + 1) + j;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int j = 0; j < CHAINS; ++j) x[j] = fma(x[j], a, b);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < CHAINS; ++j) s += x[j];
    if (s == 123456.789) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;  // never true: keeps the chains alive
}

}  // namespace

// -> TFLOP/s of fp64 FMA (2 flop each) sustained by the whole GPU; best of `reps` launches timed with CUDA events
extern "C" int ahmc_mb_dfma_peak(int device, int iters, int reps, double* tflops_out, double* ms_out) {
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    cudaDeviceProp prop{};
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1;
    const int blocks = prop.multiProcessorCount * 8, threads = 256;
    constexpr int CH = 8;
    double* sink = nullptr;
    if (cudaMalloc(&sink, sizeof(double) * blocks * threads) != cudaSuccess) return -2;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    dfma_kernel<CH><<<blocks, threads>>>(sink, iters, 0.999999, 1e-7);  // warm-up
    double best = 1e30;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        dfma_kernel<CH><<<blocks, threads>>>(sink, iters, 0.999999, 1e-7);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) return -3;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    const double flop = 2.0 * (double)blocks * threads * CH * 8.0 * iters;
    if (tflops_out) *tflops_out = flop / (best * 1e-3) / 1e12;
    if (ms_out) *ms_out = best;
    return cudaGetLastError() == cudaSuccess ? 0 : -4;
}
