// Fused moments kernel: Monomial variants (moments_kernel.cuh).
#include "moments_kernel.cuh"

namespace mlmcb200 {
namespace detail {

int launch_moments_monomial(const MomentsArgs& a, const Plan& p, bool coarse, bool is_log, cudaStream_t st) {
    if (is_log)
        return coarse ? launch_moments_s<MLMCB200_MONOMIAL, true, true>(a, p, st)
                      : launch_moments_s<MLMCB200_MONOMIAL, false, true>(a, p, st);
    return coarse ? launch_moments_s<MLMCB200_MONOMIAL, true, false>(a, p, st)
                  : launch_moments_s<MLMCB200_MONOMIAL, false, false>(a, p, st);
}

}  // namespace detail
}  // namespace mlmcb200
