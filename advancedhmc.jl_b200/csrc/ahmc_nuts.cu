// ahmc_nuts.cu -- K3 placeholder (real kernel lands next); keeps the ABI symbol set complete.
#include "ahmc_kernels.cuh"
namespace ahmc {
long long nuts_scratch_doubles_per_chain(int D, int max_depth) { return (long long)(7 + 5 * (max_depth > 0 ? max_depth : 1)) * D; }
cudaError_t launch_nuts(const NutsArgs&, cudaStream_t, int*) { return cudaErrorNotSupported; }
}  // namespace ahmc
