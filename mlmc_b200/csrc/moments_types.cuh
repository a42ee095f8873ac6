// Types shared by the translation units of the fused moments kernels (moments.cu: planning, reductions, C entry points;
// moments_k_*.cu: the kernel instantiations of one basis family each -- nvcc compiles a translation unit on one core, and
// the ~110 variants of the kernel in a single file took three minutes).
#pragma once
#include "common.cuh"

namespace mlmcb200 {
namespace detail {

constexpr int kThreads = 128;

struct MomentsArgs {
    mlmcb200_basis_t basis;
    const double* pairs;
    const uint8_t* valid;
    int64_t n;
    int64_t stride_n, stride_side, stride_m;
    int32_t n_comp;
    int32_t vec2;          // 1: scalar quantity in storage order, (fine, coarse) read as one 16-byte load
    double* partial;       // [gridDim.z][gridDim.y][2 + 2K]
    int64_t partial_stride;
    const int32_t* idx;    // re-sampling (bootstrap): replicate z reads the rows idx[z * n + i], i < n; else NULL
    int32_t fuse_mask;     // vector quantity whose components all live in this CTA, no `valid` given: the kernel combines
                           // the components' domain tests per sample itself (shared-memory flags, two barriers per tile)
};

struct Plan {
    dim3 grid;
    size_t smem;
    int S;        // samples per thread and tile
    bool pair;    // lane pairs share accumulator columns (scalar quantity)
    bool fast;    // scalar quantity in storage order: specialised addressing
    bool stream;  // few moments (HBM-bound): tiles staged through a TMA-fed shared-memory ring
    bool gather;  // re-sampled rows (bootstrap replicates in grid.z)
    bool nosq;    // sums only (no sums of squares): scalar FAST variant, private columns of half the size
};

constexpr int kStages = 4;
constexpr int kStreamMaxMoments = 12;

// Sample partitions (grid.y) of a launch with `columns` CTA columns (component blocks x bootstrap replicates): chosen so
// that columns x partitions fills whole resident waves of `slots` CTAs -- a 1.07-wave grid runs at half speed.
constexpr unsigned kMaxWavesResampled = 16;
constexpr unsigned kMaxWavesVector = 2;
unsigned choose_partitions(unsigned slots, unsigned columns, int64_t tiles, unsigned max_waves);

// One launch of the fused kernel (the variant the plan selects); returns the number of per-CTA partials written per
// batch entry (grid.y), <= 0 on error.  Defined in moments_k_*.cu.
int launch_moments_raw(const MomentsArgs& a, const Plan& p, bool coarse, cudaStream_t st);
int launch_moments_legendre_coarse(const MomentsArgs& a, const Plan& p, bool is_log, cudaStream_t st);
int launch_moments_legendre_level0(const MomentsArgs& a, const Plan& p, bool is_log, cudaStream_t st);
int launch_moments_monomial(const MomentsArgs& a, const Plan& p, bool coarse, bool is_log, cudaStream_t st);
int launch_moments_fourier(const MomentsArgs& a, const Plan& p, bool coarse, bool is_log, cudaStream_t st);

}  // namespace detail
}  // namespace mlmcb200
