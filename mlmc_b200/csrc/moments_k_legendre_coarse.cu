// Fused moments kernel: Legendre variants of levels with a coarse part (moments_kernel.cuh).
#include "moments_kernel.cuh"

namespace mlmcb200 {
namespace detail {

int launch_moments_legendre_coarse(const MomentsArgs& a, const Plan& p, bool is_log, cudaStream_t st) {
    return is_log ? launch_moments_s<MLMCB200_LEGENDRE, true, true>(a, p, st)
                  : launch_moments_s<MLMCB200_LEGENDRE, true, false>(a, p, st);
}

}  // namespace detail
}  // namespace mlmcb200
