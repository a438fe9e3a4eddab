// ahmc_pooled.cu -- the path's one exchange, behind the C ABI and on the device (DESIGN "K5c"):
// pooled (all chains of all ranks) step-size and mass-matrix adaptation with ZERO host synchronisation per warm-up
// iteration.  Per iteration, on the context stream:
//     K5 (ahmc_adapt.cu)      this rank's record  [N, sum min(1,alpha), mean_c theta, sum_c (theta - mean)^2]
//     ncclAllGather           (2 + 2D) doubles per rank -- latency only; skipped for a single rank
//     pooled_update_kernel    rank-ordered Chan merge (bit-identical on every rank), then the reference's adaptor
//                             arithmetic on the merged record: NesterovDualAveraging `adapt!` / `reset!` / `finalize!`
//                             (src/adaptation/stepsize.jl:38-62, 178-210), WelfordVar `push!` / `get_estimation`
//                             (massmatrix.jl:141-157, n_min :60-62), the Stan window schedule
//                             (stan_adaptor.jl:13-50, 137-159); eps and M^-1 are written straight into the device buffers
//                             the next transition reads (eps_chain[N], Minv[D]).
// NCCL is bound at run time (dlopen): the library loads without it and a Julia host can hand over the communicator it
// already owns (NCCL.jl) or let ahmc_comm_create make one from a broadcast unique id.
#ifndef AHMC_SIMT_EMULATION
#include <dlfcn.h>
#endif

#include <cstdio>
#include <cstring>
#include <mutex>

#include "ahmc_kernels.cuh"

namespace ahmc {

// ---------------------------------------------------------------------------------------------- device state
struct PooledState {
    // dual averaging (stepsize.jl:25-36)
    double eps, mu, x_bar, H_bar;
    double delta, gamma, t0, kappa;
    int m;
    // Welford over chains x iterations of the current window (massmatrix.jl:86-101)
    double n;
    // schedule (stan_adaptor.jl:61-72)
    int i, n_adapts, window_start, window_end, n_splits;
    int splits[16];
    int adapt_metric, n_min;
    int finalized;
};

// one CTA.  gathered: R records of (2 + 2D) doubles in rank order.
__global__ void __launch_bounds__(256) pooled_update_kernel(PooledState* st, const double* __restrict__ gathered, int R, int D,
                                                            double* w_mu, double* w_M2, double* Minv, double* eps_chain,
                                                            long long N, double* eps_trace, double* merged_out) {
    __shared__ double s_n, s_alpha;
    __shared__ int s_push, s_update, s_reset;
    const int rec = 2 + 2 * D;
    // ---- rank-ordered Chan merge of the R records (adaptation.py merge_records; fixed order => identical on all ranks)
    if (threadIdx.x == 0) {
        double n = gathered[0], a = gathered[1];
        for (int r = 1; r < R; ++r) {
            n += gathered[(size_t)r * rec];
            a = a + gathered[(size_t)r * rec + 1];
        }
        s_n = n;
        s_alpha = a;
        if (merged_out) {
            merged_out[0] = n;
            merged_out[1] = a;
        }
    }
    // schedule decisions of THIS iteration (stan_adaptor.jl:137-159)
    if (threadIdx.x == 0) {
        const int i = st->i + 1;
        bool split = false;
        for (int k = 0; k < st->n_splits; ++k) split |= (st->splits[k] == i);
        s_push = st->adapt_metric && i >= st->window_start && i <= st->window_end;
        s_update = s_push && split;
        s_reset = split;
    }
    __syncthreads();
    const bool push = s_push != 0, update = s_update != 0, reset = s_reset != 0;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        double n_a = gathered[0], mean = gathered[2 + d], M2 = gathered[2 + D + d];
        for (int r = 1; r < R; ++r) {
            const double* g = gathered + (size_t)r * rec;
            const double n_b = g[0], n = n_a + n_b, w = n_a * n_b / n;
            const double dl = g[2 + d] - mean;
            M2 += g[2 + D + d] + dl * dl * w;
            mean += dl * (n_b / n);
            n_a = n;
        }
        if (merged_out) {
            merged_out[2 + d] = mean;
            merged_out[2 + D + d] = M2;
        }
        if (push) {  // WelfordVar.push_record: Chan merge of the iteration's pooled record into the window's accumulator
            const double na = st->n, nb = s_n, n = na + nb;
            const double dl = mean - w_mu[d];
            double M = w_M2[d] + M2 + dl * dl * (na * nb / n);
            double mu = w_mu[d] + dl * (nb / n);
            if (update && n >= (double)st->n_min)  // get_estimation (massmatrix.jl:152-157)
                Minv[d] = n / ((n + 5.0) * (n - 1.0)) * M + 1e-3 * (5.0 / (n + 5.0));
            if (reset) {
                M = 0.0;
                mu = 0.0;
            }
            w_M2[d] = M;
            w_mu[d] = mu;
        } else if (reset) {
            w_M2[d] = 0.0;
            w_mu[d] = 0.0;
        }
    }
    __syncthreads();
    __shared__ double s_eps;
    if (threadIdx.x == 0) {
        const int i = st->i + 1;
        st->i = i;
        // NesterovDualAveraging.adapt (stepsize.jl:178-210) on the pooled mean of min(1, alpha)
        const double a = s_alpha / s_n;
        const int m = st->m + 1;
        const double eta_H = 1.0 / ((double)m + st->t0);
        const double H_bar = (1.0 - eta_H) * st->H_bar + eta_H * (st->delta - a);
        const double x = st->mu - H_bar * (sqrt((double)m) / st->gamma);
        const double eta_x = pow((double)m, -st->kappa);
        const double x_bar = (1.0 - eta_x) * st->x_bar + eta_x * x;
        const double eps = exp(x);
        if (finite_d(eps)) {  // stepsize.jl:199-203: a non-finite proposal keeps the previous state
            st->m = m;
            st->eps = eps;
            st->x_bar = x_bar;
            st->H_bar = H_bar;
        }
        if (push) st->n = reset ? 0.0 : st->n + s_n;
        else if (reset) st->n = 0.0;
        if (reset) {  // reset!(ssa) (stepsize.jl:38-44): restart dual averaging around the current step size
            st->m = 0;
            st->mu = log(10.0 * st->eps);
            st->x_bar = 0.0;
            st->H_bar = 0.0;
        }
        if (i == st->n_adapts) {  // finalize! (stepsize.jl:54-57)
            st->eps = exp(st->x_bar);
            st->finalized = 1;
        }
        s_eps = st->eps;
        if (eps_trace) eps_trace[i - 1] = st->eps;
    }
    __syncthreads();
    const double e = s_eps;
    for (long long c = threadIdx.x; c < N; c += blockDim.x) eps_chain[c] = e;
}

#ifndef AHMC_SIMT_EMULATION  // host launch code and the NCCL binding (skipped by the CPU SIMT emulation harness, tests/simt_emu/)
__global__ void fill_kernel(double* p, long long n, double v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}

cudaError_t launch_pooled_update(void* state, const double* gathered, int R, int D, double* w_mu, double* w_M2, double* Minv,
                                 double* eps_chain, long long N, double* eps_trace, double* merged_out, cudaStream_t st,
                                 int* n_launches) {
    pooled_update_kernel<<<1, 256, 0, st>>>((PooledState*)state, gathered, R, D, w_mu, w_M2, Minv, eps_chain, N, eps_trace, merged_out);
    if (n_launches) *n_launches += 1;
    return cudaGetLastError();
}
cudaError_t launch_fill(double* p, long long n, double v, cudaStream_t st) {
    fill_kernel<<<(unsigned)((n + 255) / 256 < 1024 ? (n + 255) / 256 : 1024), 256, 0, st>>>(p, n, v);
    return cudaGetLastError();
}
#endif  // AHMC_SIMT_EMULATION
size_t pooled_state_bytes() { return sizeof(PooledState); }

// host image of the state for creation / read-back
void pooled_state_init(void* host_image, double eps0, const AdaptDev& sched, double delta, double gamma, double t0, double kappa,
                       int n_adapts, int adapt_metric, int n_min) {
    PooledState s{};
    s.eps = eps0;
    s.mu = log(10.0 * eps0);  // stepsize.jl:38-44
    s.x_bar = 0.0;
    s.H_bar = 0.0;
    s.delta = delta; s.gamma = gamma; s.t0 = t0; s.kappa = kappa;
    s.m = 0;
    s.n = 0.0;
    s.i = 0;
    s.n_adapts = n_adapts;
    s.window_start = sched.window_start;
    s.window_end = sched.window_end;
    s.n_splits = sched.n_splits;
    for (int k = 0; k < sched.n_splits && k < 16; ++k) s.splits[k] = sched.splits[k];
    s.adapt_metric = adapt_metric;
    s.n_min = n_min;
    s.finalized = 0;
    memcpy(host_image, &s, sizeof s);
}
void pooled_state_read(const void* host_image, double* eps, int* iteration, int* m, double* n_window) {
    PooledState s;
    memcpy(&s, host_image, sizeof s);
    if (eps) *eps = s.eps;
    if (iteration) *iteration = s.i;
    if (m) *m = s.m;
    if (n_window) *n_window = s.n;
}

#ifndef AHMC_SIMT_EMULATION
// ---------------------------------------------------------------------------------------------- NCCL (bound at run time)
namespace {
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, NcclId, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    char why[256] = {0};
};
NcclApi g_nccl;
bool g_nccl_tried = false;
std::mutex g_nccl_mutex;
}  // namespace

const char* nccl_bind() {  // nullptr on success, else a reason
    std::lock_guard<std::mutex> lock(g_nccl_mutex);
    if (g_nccl.lib) return nullptr;
    if (g_nccl_tried) return g_nccl.why;
    g_nccl_tried = true;
    const char* env = getenv("AHMC_NCCL_LIB");
    const char* names[] = {env, "libnccl.so.2", "libnccl.so", "/usr/lib/x86_64-linux-gnu/libnccl.so.2"};
    for (const char* n : names) {
        if (!n) continue;
        g_nccl.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    if (!g_nccl.lib) {
        snprintf(g_nccl.why, sizeof g_nccl.why, "libnccl.so.2 not found (set AHMC_NCCL_LIB): %s", dlerror());
        return g_nccl.why;
    }
    g_nccl.GetUniqueId = (int (*)(void*))dlsym(g_nccl.lib, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void**, int, NcclId, int))dlsym(g_nccl.lib, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void*))dlsym(g_nccl.lib, "ncclCommDestroy");
    g_nccl.AllGather = (int (*)(const void*, void*, size_t, int, void*, cudaStream_t))dlsym(g_nccl.lib, "ncclAllGather");
    g_nccl.GetErrorString = (const char* (*)(int))dlsym(g_nccl.lib, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather) {
        snprintf(g_nccl.why, sizeof g_nccl.why, "NCCL library lacks ncclGetUniqueId / ncclCommInitRank / ncclAllGather");
        g_nccl.lib = nullptr;
        return g_nccl.why;
    }
    return nullptr;
}
const char* nccl_err(int rc) { return g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "nccl error"; }
int nccl_unique_id(void* out128) { return g_nccl.GetUniqueId(out128); }
int nccl_comm_init(void** comm, int nranks, const void* id128, int rank) {
    NcclId id;
    memcpy(id.internal, id128, 128);
    return g_nccl.CommInitRank(comm, nranks, id, rank);
}
int nccl_comm_destroy(void* comm) { return g_nccl.CommDestroy(comm); }
int nccl_allgather_f64(const double* send, double* recv, size_t count, void* comm, cudaStream_t st) {
    return g_nccl.AllGather(send, recv, count, 8 /* ncclFloat64 */, comm, st);
}
#endif  // AHMC_SIMT_EMULATION

}  // namespace ahmc
