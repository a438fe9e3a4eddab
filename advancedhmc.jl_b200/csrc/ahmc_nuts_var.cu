// ahmc_nuts_var.cu -- K3, variant family: SliceTS sampler, ClassicNoUTurn / StrictGeneralisedNoUTurn criteria
// (src/trajectory.jl:102-109, 551-557, 579-613).  Kernel: ahmc_nuts_kernel.cuh.
#include "ahmc_nuts_kernel.cuh"

namespace ahmc {

cudaError_t launch_nuts_variants(const NutsArgs& a, cudaStream_t st) { return nuts_dispatch<true, false, false>(a, st); }

}  // namespace ahmc
