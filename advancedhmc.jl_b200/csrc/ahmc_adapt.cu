// ahmc_adapt.cu -- K5: pooled adaptor statistics of one iteration over this GPU's chains.
//
// The reference adapts per chain and never pools (src/adaptation/Adaptation.jl:52 "TODO: implement
// consensus adaptor"); many-chain HMCDA/NUTS with a SHARED step size and metric needs the pooled form
// (SURVEY.md section 8a Q6, 8e).  This kernel produces the mergeable record
//     out[0] = N                      out[1] = sum_c min(1, alpha_c)          (stepsize.jl:178-210 input)
//     out[2 .. 2+D)   = mean_c theta[:, c]
//     out[2+D .. 2+2D) = sum_c (theta[:, c] - mean)^2                          (Welford (n, mu, M2),
//                                                                              massmatrix.jl:141-149)
// that the host all-gathers across ranks and Chan-merges in rank order.  Deterministic: fixed grid,
// per-block partials, the last block to finish reduces them in block order.
#include "ahmc_kernels.cuh"

namespace ahmc {

constexpr int kAdaptThreads = 256;

// pass 0: column sums (+ alpha sum); pass 1: centred squares given the mean already in out[2..2+D)
template <int PASS>
__global__ void __launch_bounds__(kAdaptThreads) adapt_kernel(int D, long long N, const double* __restrict__ theta,
                                                              long long ld, const double* __restrict__ alpha,
                                                              double* partial /* gridDim.x * (D+1) */,
                                                              unsigned* counter, double* out) {
    const int B = gridDim.x;
    const long long per = (N + B - 1) / B;
    const long long c0 = per * blockIdx.x, c1 = (c0 + per < N) ? c0 + per : N;
    // thread t owns coordinates d = t, t+T, ...; loops over this block's chains (coalesced along d)
    for (int d = threadIdx.x; d < D; d += kAdaptThreads) {
        double acc = 0.0;
        const double mu = PASS == 1 ? out[2 + d] : 0.0;
        for (long long c = c0; c < c1; ++c) {
            double v = theta[ld * c + d];
            if (PASS == 0) acc += v;
            else acc = fma(v - mu, v - mu, acc);
        }
        partial[(size_t)blockIdx.x * (D + 1) + d] = acc;
    }
    if (PASS == 0 && threadIdx.x == 0) {
        double a = 0.0;
        if (alpha)
            for (long long c = c0; c < c1; ++c) {
                double x = alpha[c];
                a += (x < 1.0) ? x : ((x != x) ? x : 1.0);  // min(1, alpha), NaN-propagating like Julia
            }
        partial[(size_t)blockIdx.x * (D + 1) + D] = a;
    }
    __threadfence();
    __shared__ bool last;
    __syncthreads();
    if (threadIdx.x == 0) last = (atomicAdd(counter, 1u) == (unsigned)(B - 1));
    __syncthreads();
    if (!last) return;
    __threadfence();
    for (int d = threadIdx.x; d < D; d += kAdaptThreads) {
        double s = 0.0;
        for (int b = 0; b < B; ++b) s += partial[(size_t)b * (D + 1) + d];
        if (PASS == 0) out[2 + d] = s / (double)N;
        else out[2 + D + d] = s;
    }
    if (PASS == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int b = 0; b < B; ++b) s += partial[(size_t)b * (D + 1) + D];
        out[0] = (double)N;
        out[1] = s;
    }
    if (threadIdx.x == 0) *counter = 0u;  // re-arm for the next launch
}

#ifndef AHMC_SIMT_EMULATION  // host launch code (skipped by the CPU SIMT emulation harness, tests/simt_emu/)
cudaError_t launch_adapt_summary(int D, long long N, const double* theta, long long ld, const double* alpha,
                                 double* out, double* partial, unsigned* counter, int blocks, cudaStream_t st,
                                 int* n_launches) {
    adapt_kernel<0><<<blocks, kAdaptThreads, 0, st>>>(D, N, theta, ld, alpha, partial, counter, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    adapt_kernel<1><<<blocks, kAdaptThreads, 0, st>>>(D, N, theta, ld, alpha, partial, counter, out);
    if (n_launches) *n_launches += 2;
    return cudaGetLastError();
}

#endif  // AHMC_SIMT_EMULATION

// K5b: full second-moment matrix about a given mean, for the pooled WelfordCov (massmatrix.jl:286-340):
//     out[i + D*j] = sum_c (theta[i, c] - mean[i]) * (theta[j, c] - mean[j])            (symmetric, D x D)
// One CTA per 32x32 output tile of the upper triangle walks ALL chains in slabs of 32 (centred slabs staged in
// shared memory, 2x2 outputs per thread), so every entry is one fixed-order sum: deterministic, no atomics; the
// mirror entry is written by the same thread.
constexpr int kCovTile = 32;

__global__ void __launch_bounds__(256) adapt_cov_kernel(int D, long long N, const double* __restrict__ theta, long long ld,
                                                         const double* __restrict__ mean, double* __restrict__ out) {
    // tile index -> (ti, tj) with ti <= tj
    const int T = (D + kCovTile - 1) / kCovTile;
    int t = blockIdx.x, ti = 0;
    while (t >= T - ti) {
        t -= T - ti;
        ++ti;
    }
    const int tj = ti + t;
    __shared__ double A[kCovTile][kCovTile + 1];  // [chain in slab][coordinate i]
    __shared__ double Bm[kCovTile][kCovTile + 1];
    const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
    const int li = threadIdx.x % kCovTile, lc = threadIdx.x / kCovTile;  // loader mapping: 8 chains x 32 coordinates
    const int gi = ti * kCovTile + li, gj = tj * kCovTile + li;
    const double mi = gi < D ? mean[gi] : 0.0, mj = gj < D ? mean[gj] : 0.0;
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
    for (long long c0 = 0; c0 < N; c0 += kCovTile) {
#pragma unroll
        for (int k = 0; k < kCovTile / 8; ++k) {
            const int cc = lc + 8 * k;
            const long long c = c0 + cc;
            const bool ok = c < N;
            A[cc][li] = (ok && gi < D) ? theta[ld * c + gi] - mi : 0.0;
            Bm[cc][li] = (ok && gj < D) ? theta[ld * c + gj] - mj : 0.0;
        }
        __syncthreads();
#pragma unroll 8
        for (int cc = 0; cc < kCovTile; ++cc) {
            const double a0 = A[cc][ty], a1 = A[cc][ty + 16], b0 = Bm[cc][tx], b1 = Bm[cc][tx + 16];
            acc[0][0] = fma(a0, b0, acc[0][0]);
            acc[0][1] = fma(a0, b1, acc[0][1]);
            acc[1][0] = fma(a1, b0, acc[1][0]);
            acc[1][1] = fma(a1, b1, acc[1][1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int i = ti * kCovTile + ty + 16 * p, j = tj * kCovTile + tx + 16 * q;
            if (i < D && j < D) {
                out[(size_t)i + (size_t)D * j] = acc[p][q];
                if (ti != tj) out[(size_t)j + (size_t)D * i] = acc[p][q];
            }
        }
}

#ifndef AHMC_SIMT_EMULATION
cudaError_t launch_adapt_cov(int D, long long N, const double* theta, long long ld, const double* mean, double* out,
                             cudaStream_t st, int* n_launches) {
    const int T = (D + kCovTile - 1) / kCovTile;
    adapt_cov_kernel<<<T * (T + 1) / 2, 256, 0, st>>>(D, N, theta, ld, mean, out);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}

#endif  // AHMC_SIMT_EMULATION

}  // namespace ahmc
