// Fused moments kernel: RAW and Fourier variants (moments_kernel.cuh).
#include "moments_kernel.cuh"

namespace mlmcb200 {
namespace detail {

int launch_moments_raw(const MomentsArgs& a, const Plan& p, bool coarse, cudaStream_t st) {
    return coarse ? launch_moments_s<MLMCB200_RAW, true, false>(a, p, st)
                  : launch_moments_s<MLMCB200_RAW, false, false>(a, p, st);
}

int launch_moments_fourier(const MomentsArgs& a, const Plan& p, bool coarse, bool is_log, cudaStream_t st) {
    if (is_log)
        return coarse ? launch_moments_s<MLMCB200_FOURIER, true, true>(a, p, st)
                      : launch_moments_s<MLMCB200_FOURIER, false, true>(a, p, st);
    return coarse ? launch_moments_s<MLMCB200_FOURIER, true, false>(a, p, st)
                  : launch_moments_s<MLMCB200_FOURIER, false, false>(a, p, st);
}

}  // namespace detail
}  // namespace mlmcb200
