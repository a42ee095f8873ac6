// Exact order statistics of a column of doubles by radix selection (sm_100a).
//
// Replaces the `np.percentile(fine_samples, [100 q, 100 (1 - q)])` of Estimate.estimate_domain
// (mlmc/estimator.py:275-302): numpy sorts (partitions) the whole column to read two pairs of neighbouring order
// statistics; here the same four values are found by six histogram passes over the data (11 + 11 + 11 + 11 + 11 + 9 key
// bits), no sort, no copy.  The values are the exact order statistics, so the percentile the host interpolates from
// them is bit-identical to numpy's.
//
// A double maps to a uint64 key that orders like the number (sign flip); NaN entries are skipped (the reference drops
// masked samples before the percentile).  All requested ranks are selected together: in every pass an element is
// counted in the histogram of each rank whose key prefix it still matches (per-CTA shared-memory histograms, one
// global atomic per non-empty bin), and a one-CTA kernel turns the counts into the next key digit per rank.
#include "common.cuh"

namespace mlmcb200 {
namespace {

constexpr int kSelBits = 11, kSelBins = 1 << kSelBits, kSelPasses = 6, kSelThreads = 256;
constexpr int kMaxFracs = 8, kMaxRanks = 2 * kMaxFracs;

struct SelectState {
    unsigned long long prefix[kMaxRanks];     // key bits decided so far (high bits), per rank
    long long rank[kMaxRanks];                // rank among the elements that match the prefix
    long long n_valid;
};

struct SelectArgs {
    const double* x;
    int64_t n, stride;
    int n_ranks;
    double frac[kMaxFracs];
    SelectState* state;
    unsigned long long* hist;                 // [n_ranks][kSelBins]
    double* out;                              // [n_ranks] order statistics, then n_valid
};

__host__ __device__ inline int pass_shift(int pass) { return pass < 5 ? 64 - kSelBits * (pass + 1) : 0; }
__host__ __device__ inline int pass_bits(int pass) { return pass < 5 ? kSelBits : 64 - 5 * kSelBits; }

__device__ __forceinline__ unsigned long long key_of(double v) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

__device__ __forceinline__ double value_of(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(kSelThreads) select_hist_kernel(const SelectArgs a, int pass) {
    __shared__ unsigned hist_s[4 * kSelBins];                 // the histograms of 4 ranks at a time (32 KB)
    const int shift = pass_shift(pass), bits = pass_bits(pass);
    const unsigned digit_mask = (1u << bits) - 1u;
    // pass 0: no prefix yet, every rank sees the same histogram -> counted once, into the histogram of rank 0
    const int n_ranks = pass == 0 ? 1 : a.n_ranks;
    for (int r0 = 0; r0 < n_ranks; r0 += 4) {
        const int nr = min(4, n_ranks - r0);
        for (int i = threadIdx.x; i < nr * kSelBins; i += kSelThreads) hist_s[i] = 0;
        __syncthreads();
        unsigned long long pre[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) pre[r] = r < nr ? a.state->prefix[r0 + r] : 0ull;
        for (int64_t i = (int64_t)blockIdx.x * kSelThreads + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * kSelThreads) {
            const double v = __ldg(a.x + i * a.stride);
            if (v != v) continue;
            const unsigned long long k = key_of(v);
            const unsigned long long hi = pass == 0 ? 0ull : (k >> (shift + bits));
            const unsigned d = (unsigned)(k >> shift) & digit_mask;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (r < nr && (pass == 0 || hi == (pre[r] >> (shift + bits)))) atomicAdd(&hist_s[r * kSelBins + d], 1u);
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < nr * kSelBins; i += kSelThreads) {
            const unsigned c = hist_s[i];
            if (c) atomicAdd(&a.hist[(size_t)r0 * kSelBins + i], (unsigned long long)c);
        }
        __syncthreads();
    }
}

// one CTA: per rank, the digit whose cumulative count crosses the rank; clears the histogram for the next pass
__global__ void __launch_bounds__(kSelThreads) select_pick_kernel(const SelectArgs a, int pass) {
    __shared__ unsigned long long part[kSelThreads];
    const int shift = pass_shift(pass), bins = 1 << pass_bits(pass);
    const int per = kSelBins / kSelThreads;                   // bins per thread (8)
    for (int r = 0; r < a.n_ranks; ++r) {
        const unsigned long long* h = a.hist + (size_t)(pass == 0 ? 0 : r) * kSelBins;
        unsigned long long mine = 0;
        for (int j = 0; j < per; ++j) {
            const int b = threadIdx.x * per + j;
            if (b < bins) mine += h[b];
        }
        part[threadIdx.x] = mine;
        __syncthreads();
        if (threadIdx.x == 0) {
            if (pass == 0) {
                // ranks from the fractions, with numpy's arithmetic: pos = frac * (n - 1); lo = floor(pos); hi = lo + 1
                unsigned long long total = 0;
                for (int t = 0; t < kSelThreads; ++t) total += part[t];
                a.state->n_valid = (long long)total;
                const double pos = __dmul_rn(a.frac[r >> 1], (double)((long long)total - 1));
                long long lo = (long long)floor(pos);
                if (lo < 0) lo = 0;
                long long hi = lo + 1 < (long long)total ? lo + 1 : (long long)total - 1;
                if (hi < 0) hi = 0;
                a.state->rank[r] = (r & 1) ? hi : lo;
                a.state->prefix[r] = 0ull;
            }
            long long want = a.state->rank[r];
            unsigned long long cum = 0;
            int t = 0;
            for (; t < kSelThreads - 1; ++t) {
                if (cum + part[t] > (unsigned long long)want) break;
                cum += part[t];
            }
            int b = t * per;
            for (; b < t * per + per - 1 && b < bins - 1; ++b) {
                if (cum + h[b] > (unsigned long long)want) break;
                cum += h[b];
            }
            a.state->prefix[r] |= (unsigned long long)b << shift;
            a.state->rank[r] = want - (long long)cum;
            if (pass == kSelPasses - 1) a.out[r] = value_of(a.state->prefix[r]);
        }
        __syncthreads();
    }
    for (int i = threadIdx.x; i < a.n_ranks * kSelBins; i += kSelThreads) a.hist[i] = 0ull;
    if (pass == kSelPasses - 1 && threadIdx.x == 0) a.out[a.n_ranks] = (double)a.state->n_valid;
}

}  // namespace
}  // namespace mlmcb200

using namespace mlmcb200;

extern "C" int64_t mlmcb200_percentile_workspace_bytes(int32_t n_frac) {
    if (n_frac < 1 || n_frac > kMaxFracs) return -1;
    return (int64_t)sizeof(SelectState) + (int64_t)2 * n_frac * kSelBins * (int64_t)sizeof(unsigned long long);
}

extern "C" int mlmcb200_percentile_stats(const double* x, int64_t n, int64_t stride, const double* frac,
                                         int32_t n_frac, double* out, void* workspace, int64_t workspace_bytes,
                                         void* stream) {
    MB_REQUIRE(x != nullptr && out != nullptr && workspace != nullptr && frac != nullptr, "percentile_stats: null pointer");
    MB_REQUIRE(n >= 1 && stride >= 1, "percentile_stats: bad n=%lld stride=%lld", (long long)n, (long long)stride);
    MB_REQUIRE(n_frac >= 1 && n_frac <= kMaxFracs, "percentile_stats: 1..%d fractions, got %d", kMaxFracs, n_frac);
    MB_REQUIRE(workspace_bytes >= mlmcb200_percentile_workspace_bytes(n_frac), "percentile_stats: workspace too small");
    for (int i = 0; i < n_frac; ++i)
        MB_REQUIRE(frac[i] >= 0.0 && frac[i] <= 1.0, "percentile_stats: fraction %g outside [0, 1]", frac[i]);
    cudaStream_t st = (cudaStream_t)stream;
    SelectArgs a;
    a.x = x;
    a.n = n;
    a.stride = stride;
    a.n_ranks = 2 * n_frac;
    for (int i = 0; i < kMaxFracs; ++i) a.frac[i] = i < n_frac ? frac[i] : 0.0;
    a.state = static_cast<SelectState*>(workspace);
    a.hist = reinterpret_cast<unsigned long long*>(static_cast<char*>(workspace) + sizeof(SelectState));
    a.out = out;
    MB_CUDA_OK(cudaMemsetAsync(workspace, 0, (size_t)mlmcb200_percentile_workspace_bytes(n_frac), st));
    int64_t blocks = (n + kSelThreads * 8 - 1) / (kSelThreads * 8);
    const int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    for (int pass = 0; pass < kSelPasses; ++pass) {
        select_hist_kernel<<<(unsigned)blocks, kSelThreads, 0, st>>>(a, pass);
        MB_CUDA_OK(cudaGetLastError());
        select_pick_kernel<<<1, kSelThreads, 0, st>>>(a, pass);
        MB_CUDA_OK(cudaGetLastError());
    }
    return 0;
}
