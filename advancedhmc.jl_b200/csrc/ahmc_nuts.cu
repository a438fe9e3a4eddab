// ahmc_nuts.cu -- K3 entry: workspace size, family dispatch, and the default family
// (MultinomialTS + GeneralisedNoUTurn, what `NUTS(delta)` builds).  Kernel: ahmc_nuts_kernel.cuh.
#include "ahmc_nuts_kernel.cuh"

namespace ahmc {

long long nuts_scratch_doubles_per_chain(int D, int max_depth, bool adaptive) {
    return nuts_level_doubles(D, max_depth) + (adaptive ? 2LL * D : 0);
}

cudaError_t launch_nuts_variants(const NutsArgs& a, cudaStream_t st);  // ahmc_nuts_var.cu
cudaError_t launch_nuts_adaptive(const NutsArgs& a, cudaStream_t st);  // ahmc_nuts_adapt.cu

cudaError_t launch_nuts(const NutsArgs& a, cudaStream_t st, int* n_launches) {
    if (n_launches) *n_launches += 1;
    if (a.ad.enabled) return launch_nuts_adaptive(a, st);
    if (a.sampler != 0 || a.criterion != 0) return launch_nuts_variants(a, st);
    return nuts_dispatch<false, false, false>(a, st);
}

}  // namespace ahmc
