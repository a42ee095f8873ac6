// Fused moments path: planning, partial reductions, mask / finalize / transform kernels and the C entry points.
// The accumulate kernel itself (moments_kernel.cuh) is instantiated per basis family in moments_k_*.cu.
#include <stdlib.h>
#include "moments_types.cuh"
#include "philox.cuh"

namespace mlmcb200 {

namespace {

using namespace detail;

// acc[j] += sum_b partial[b][j] in a fixed order (bitwise reproducible).  One warp per output when there are many
// partials (lanes take b = lane, lane + 32, ... then a shuffle tree), one thread per output otherwise.
// blockIdx.y = batch entry (bootstrap replicate): its partials follow those of the previous entry, its accumulator is
// acc_batch_stride doubles further.
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int n_partials, int64_t stride,
                                       int64_t len, double* __restrict__ acc, int warp_per_output,
                                       int64_t acc_batch_stride) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    partial += (int64_t)blockIdx.y * n_partials * stride;
    acc += (int64_t)blockIdx.y * acc_batch_stride;
    if (warp_per_output) {
        const int64_t j = gid >> 5;
        const int lane = threadIdx.x & 31;
        if (j >= len) return;
        // the loads of a lane are independent: batches of 12 are issued together (one L2 round trip per batch instead
        // of one per partial -- a one-wave launch has ~300 partials), then added in index order
        double s = 0.0;
        for (int b0 = lane; b0 < n_partials; b0 += 32 * 12) {
            double v[12];
#pragma unroll
            for (int u = 0; u < 12; ++u) {
                const int b = b0 + 32 * u;
                v[u] = b < n_partials ? partial[(int64_t)b * stride + j] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 12; ++u) s += v[u];
        }
        s = warp_sum(s);
        if (lane == 0) acc[j] += s;
    } else {
        if (gid >= len) return;
        double s = 0.0;
        for (int b = 0; b < n_partials; ++b) s += partial[(int64_t)b * stride + gid];
        acc[gid] += s;
    }
}

// ---- row numbers of the bootstrap replicates (counter-based Philox4x32-10, philox.cuh: reproducible, no state) ----
// Replicate blockIdx.y, draws 2*g and 2*g+1 of thread g.  Draw j belongs to the row block p with
// cum[p] <= j < cum[p+1] (block p = rows [p n / P, (p+1) n / P)) and is uniform inside it: floor(r64 * size / 2^64).
// The Philox counter carries the GLOBAL replicate number (rep_offset + blockIdx.y) and the key only the seed and the
// stream id, so a replicate's draws do not depend on how the replicates are grouped into launches or sharded over ranks.
__global__ void resample_indices_kernel(uint64_t seed, uint64_t stream_id, int64_t n_rows, int64_t n_draws,
                                        int n_blocks, const int64_t* __restrict__ block_cum, int32_t* __restrict__ idx,
                                        uint32_t rep_offset) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int rep = blockIdx.y;
    if (2 * g >= n_draws) return;
    uint32_t c[4] = {(uint32_t)g, (uint32_t)(g >> 32), rep_offset + (uint32_t)rep, (uint32_t)stream_id};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(stream_id >> 32));
    const int64_t* cum = block_cum ? block_cum + (int64_t)rep * (n_blocks + 1) : nullptr;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int64_t j = 2 * g + h;
        if (j >= n_draws) break;
        int64_t lo = 0, size = n_rows;
        if (cum != nullptr) {
            int a = 0, b = n_blocks;                 // largest p with cum[p] <= j
            while (b - a > 1) {
                const int mid = (a + b) >> 1;
                if (__ldg(cum + mid) <= j) a = mid;
                else b = mid;
            }
            lo = (int64_t)a * n_rows / n_blocks;
            size = (int64_t)(a + 1) * n_rows / n_blocks - lo;
        }
        const uint64_t r = ((uint64_t)c[2 * h + 1] << 32) | c[2 * h];
        idx[(int64_t)rep * n_draws + j] = (int32_t)(lo + (int64_t)__umul64hi(r, (uint64_t)size));
    }
}

// valid[] starts at 1; every (sample, side, component) value is visited once, flat and coalesced, 8 independent loads
// in flight per thread, and a value outside the domain clears its sample's flag (all writers store the same 0: no
// atomics, no per-sample reduction).  HBM-bound for any number of components.
// IDX: uint32_t when the number of values fits (one 32-bit division per value; a 64-bit division costs ~100 instructions
// and made the pass compute-bound), uint64_t otherwise.
template <typename IDX>
__global__ void sample_mask_kernel(const mlmcb200_basis_t basis, const double* __restrict__ pairs, int64_t n,
                                   int n_comp, int64_t stride_n, int64_t stride_side, int64_t stride_m,
                                   int n_sides, uint8_t* __restrict__ valid) {
    constexpr int kUnroll = 8;
    const int64_t per = (int64_t)n_sides * n_comp, total = n * per;
    const IDX per_i = (IDX)per;
    const int64_t step = (int64_t)gridDim.x * blockDim.x * kUnroll;
    for (int64_t base = (int64_t)blockIdx.x * blockDim.x * kUnroll + threadIdx.x; base < total; base += step) {
        double x[kUnroll];
        int64_t smp[kUnroll];
        // (sample, offset in the sample) of the first value by ONE division, of the following ones incrementally
        IDX si = (IDX)base / per_i;
        uint32_t r = (uint32_t)((IDX)base - si * per_i);
        int64_t s = (int64_t)si;
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const int64_t e = base + (int64_t)u * blockDim.x;
            smp[u] = -1;
            x[u] = 0.0;
            if (e < total) {
                const int side = r >= (uint32_t)n_comp ? 1 : 0, mm = (int)r - side * n_comp;      // n_sides <= 2
                smp[u] = s;
                x[u] = __ldg(pairs + s * stride_n + side * stride_side + (int64_t)mm * stride_m);
            }
            r += blockDim.x;
            if (r >= (uint32_t)per) {
                if (per >= (int64_t)blockDim.x) {
                    r -= (uint32_t)per;
                    ++s;
                } else {
                    const uint32_t q = r / (uint32_t)per;
                    r -= q * (uint32_t)per;
                    s += q;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            if (smp[u] >= 0) {
                const double t = basis.kind == MLMCB200_RAW ? x[u] : map_to_ref(basis, x[u]);
                if (!moments_finite(basis, t)) valid[smp[u]] = 0;
            }
        }
    }
}

// The same for the storage layout (the fine and coarse values of all components of a sample form ONE contiguous,
// 16-byte aligned run of 2 * per2 doubles; rows stride_n apart): the unit of work is (sample, segment of kSegment
// 16-byte words), one warp per unit -- the lanes stream the segment with 16-byte loads, four in flight per lane, and
// combine their verdicts by a vote.  No per-element index division, no scattered flag stores, twice the bytes per load
// instruction, and the parallelism does not depend on the number of samples (cfg5: 256 samples of 20 000 values).
constexpr int kMaskSegment = 512;
__global__ void sample_mask_rows_kernel(const mlmcb200_basis_t basis, const double* __restrict__ pairs, int64_t n,
                                        int per2, int64_t stride_n, uint8_t* __restrict__ valid) {
    constexpr int kUnroll = 4;
    const int lane = threadIdx.x & 31;
    const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int segs = (per2 + kMaskSegment - 1) / kMaskSegment;
    const int64_t units = n * segs;
    for (int64_t unit = warp; unit < units; unit += n_warps) {
        const int64_t s = unit / segs;
        const int j_lo = (int)(unit - s * segs) * kMaskSegment;
        const int j_hi = min(per2, j_lo + kMaskSegment);
        const double2* row = reinterpret_cast<const double2*>(pairs + s * stride_n);
        bool ok = true;
        for (int j0 = j_lo + lane; j0 < j_hi; j0 += 32 * kUnroll) {
            double2 v[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const int j = j0 + 32 * u;
                v[u] = j < j_hi ? __ldcs(row + j) : make_double2(0.0, 0.0);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                if (j0 + 32 * u < j_hi) {
                    const double ta = basis.kind == MLMCB200_RAW ? v[u].x : map_to_ref(basis, v[u].x);
                    const double tb = basis.kind == MLMCB200_RAW ? v[u].y : map_to_ref(basis, v[u].y);
                    ok = ok && moments_finite(basis, ta) && moments_finite(basis, tb);
                }
            }
        }
        ok = __all_sync(0xffffffffu, ok);
        if (lane == 0 && !ok) valid[s] = 0;                  // segments of a sample may race here: all write 0
    }
}

// blockIdx.y = batch entry (bootstrap replicate): accumulators acc_batch_stride apart, outputs packed per entry as
// [l_means (L*K) | l_vars (L*K) | mean (K) | var (K)] when out_batch_stride != 0
__global__ void finalize_levels_kernel(const double* __restrict__ acc, int64_t acc_stride, int n_levels, int64_t K,
                                       double* l_means, double* l_vars, double* mean, double* var,
                                       int64_t acc_batch_stride, int64_t out_batch_stride, double* counts = nullptr) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (counts != nullptr && k < 2 * (int64_t)n_levels)        // [n_samples, n_rm_samples] per level, packed
        counts[k] = acc[(k >> 1) * acc_stride + (k & 1)];
    if (k >= K) return;
    acc += (int64_t)blockIdx.y * acc_batch_stride;
    const int64_t out_off = (int64_t)blockIdx.y * out_batch_stride;
    if (l_means) l_means += out_off;
    if (l_vars) l_vars += out_off;
    if (mean) mean += out_off;
    if (var) var += out_off;
    double m_tot = 0.0, v_tot = 0.0;
    for (int l = 0; l < n_levels; ++l) {
        const double* a = acc + (int64_t)l * acc_stride;
        const double n = a[0];
        const double s = a[2 + k], sq = a[2 + K + k];
        // quantity_estimate.py:72-77 : s / n ; (sp - s**2 / n) / (n - 1)
        const double lm = __ddiv_rn(s, n);
        double lv;
        if (n > 1.0)
            lv = __ddiv_rn(__dsub_rn(sq, __ddiv_rn(__dmul_rn(s, s), n)), n - 1.0);
        else
            lv = __longlong_as_double(0x7ff0000000000000LL);
        if (l_means) l_means[(int64_t)l * K + k] = lm;
        if (l_vars) l_vars[(int64_t)l * K + k] = lv;
        m_tot = __dadd_rn(m_tot, lm);                 // quantity.py:592-593
        v_tot = __dadd_rn(v_tot, __ddiv_rn(lv, n));
    }
    if (mean) mean[k] = m_tot;
    if (var) var[k] = v_tot;
}

// Linear map of level sums (products of basis functions are linear combinations of a longer basis of the same family,
// so the covariance level sums are C . (level sums of 2R-1 moments)): per level l and component m
//     out[l][2 + m K1 + o] = sum_k mat_t[k K1 + o] * in[l][2 + m K0 + k]      (k ascending: reproducible)
// counts copied, the sums of squares of `out` are set to NaN (a linear map of sums says nothing about them).
__global__ void level_sums_transform_kernel(const double* __restrict__ in, int64_t in_stride, int K0, int n_comp,
                                            const double* __restrict__ mat_t, int K1, double* __restrict__ out,
                                            int64_t out_stride) {
    const int o = blockIdx.x * blockDim.x + threadIdx.x;
    const int l = blockIdx.y / n_comp, m = blockIdx.y - l * n_comp;
    const double* src = in + (int64_t)l * in_stride;
    double* dst = out + (int64_t)l * out_stride;
    if (o == 0 && m == 0) {
        dst[0] = src[0];
        dst[1] = src[1];
    }
    if (o >= K1) return;
    const double* s = src + 2 + (int64_t)m * K0;
    double v = 0.0;
    for (int k = 0; k < K0; ++k) v = fma(__ldg(mat_t + (int64_t)k * K1 + o), s[k], v);
    const int64_t K = (int64_t)n_comp * K1;
    dst[2 + (int64_t)m * K1 + o] = v;
    dst[2 + K + (int64_t)m * K1 + o] = __longlong_as_double(0x7ff8000000000000LL);
}

}  // namespace

int launch_reduce_partials_batched(const double* partial, int n_partials, int64_t stride, int64_t len, double* acc,
                                   int n_batch, int64_t acc_batch_stride, cudaStream_t st) {
    const int threads = 256;
    const int wpo = n_partials >= 32 && len <= (1 << 20) ? 1 : 0;
    const int64_t total = wpo ? len * 32 : len;
    const dim3 grid((unsigned)((total + threads - 1) / threads), (unsigned)n_batch);
    reduce_partials_kernel<<<grid, threads, 0, st>>>(partial, n_partials, stride, len, acc, wpo, acc_batch_stride);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}

int launch_reduce_partials(const double* partial, int n_partials, int64_t stride, int64_t len, double* acc,
                           cudaStream_t st) {
    return launch_reduce_partials_batched(partial, n_partials, stride, len, acc, 1, 0, st);
}

namespace {



// Samples per thread and tile: 8 fine+coarse pairs = 16 independent recurrence chains per thread (Fourier carries twice
// the state per sample and uses 4); a scalar level without coarse part takes 16 samples to keep the same 16 chains.
int choose_S(int kind, bool coarse, bool scalar, bool stream) {
    if (kind == MLMCB200_FOURIER) return 4;
    return (!coarse && scalar && kind != MLMCB200_RAW && !stream) ? 16 : 8;
}

int plan_moments(int kind, int size, int n_comp, int64_t n, bool coarse, bool stream, Plan* p, bool nosq = false) {
    p->S = choose_S(kind, coarse, n_comp == 1, stream);
    p->nosq = nosq;
    // Lane-pair columns halve the shared memory per thread at the price of a shuffle in every moment's reduction tail:
    // worth it only where private columns would cap the CTAs per SM below what the registers allow (2 for the
    // fine+coarse kernels), i.e. above ~56 moments.  MLMCB200_PAIR=0|1 overrides (experiments).
    static int forced = -2;
    if (forced == -2) {
        const char* e = getenv("MLMCB200_PAIR");
        forced = e ? atoi(e) : -1;
    }
    p->pair = n_comp == 1 && (size_t)2 * size * kThreads * sizeof(double) > 113u * 1024u;
    if (forced >= 0 && n_comp == 1) p->pair = forced != 0;
    if (nosq) p->pair = false;
    p->fast = false;
    p->gather = false;
    p->stream = stream;
    const size_t smem = (size_t)((p->pair || nosq) ? 1 : 2) * size * kThreads * sizeof(double);
    if (smem + 64 > 227u * 1024u) {        // 64 B: static shared memory of the kernel (sample counters)
        set_error("moments: size %d needs %zu B of shared memory per CTA (max %u)", size, smem, 227u * 1024u);
        return -1;
    }
    int ctas_per_sm = (int)((227u * 1024u) / (smem + 1024));
    if (ctas_per_sm > 8) ctas_per_sm = 8;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const int gx = n_comp >= kThreads ? (n_comp + kThreads - 1) / kThreads : 1;
    // gx == 1: one resident wave of sample partitions.  gx > 1 (vector quantity, one CTA column per 128 components):
    // the launch picks the partition count that fills whole waves (up to 2); this is its upper bound (workspace).
    int gy = gx == 1 ? sm_count() * ctas_per_sm : (2 * sm_count() * ctas_per_sm) / gx + 1;
    if (n >= 0) {
        const int TN = n_comp >= kThreads ? 1 : kThreads / n_comp;
        const int64_t tiles = (n + (int64_t)p->S * TN - 1) / ((int64_t)p->S * TN);
        if (tiles < gy) gy = (int)(tiles > 0 ? tiles : 1);
    }
    p->grid = dim3(gx, gy, 1);
    p->smem = smem;
    return 0;
}

}  // namespace

}  // namespace mlmcb200

namespace mlmcb200 {
namespace detail {
unsigned choose_partitions(unsigned slots, unsigned columns, int64_t tiles, unsigned max_waves) {
    unsigned best = 1;
    double best_eff = 0.0;
    for (unsigned m = 1; m <= max_waves; ++m) {
        unsigned gy = (unsigned)(((uint64_t)m * slots) / columns);
        if (gy < 1) gy = 1;
        if ((int64_t)gy * 4 > tiles && gy > 1) break;            // keep at least 4 tiles per CTA
        const uint64_t ctas = (uint64_t)gy * columns;
        const double eff = (double)ctas / (double)(((ctas + slots - 1) / slots) * slots);
        if (eff > best_eff + 0.02) {                              // a further wave must pay for its CTA start-up cost
            best_eff = eff;
            best = gy;
        }
    }
    if ((int64_t)best > tiles) best = (unsigned)(tiles > 0 ? tiles : 1);
    return best;
}
}  // namespace detail
}  // namespace mlmcb200

using namespace mlmcb200;
using namespace mlmcb200::detail;

extern "C" int64_t mlmcb200_moments_workspace_bytes(int32_t size, int32_t n_comp) {
    Plan p;
    if (size < 1 || n_comp < 1) return -1;
    if (plan_moments(MLMCB200_LEGENDRE, size, n_comp, -1, true, false, &p) != 0) return -1;
    Plan q;                                             // the sums-only variant may keep more CTAs per SM resident
    if (plan_moments(MLMCB200_LEGENDRE, size, n_comp, -1, true, false, &q, true) == 0 && q.grid.y > p.grid.y) p = q;
    return (int64_t)p.grid.y * (2 + 2 * (int64_t)size * n_comp) * (int64_t)sizeof(double);
}

namespace {

// shared body of the plain (idx == NULL, n_rep == 1) and the re-sampled accumulate calls
int moments_accumulate_impl(const mlmcb200_basis_t* basis, const double* pairs, int64_t n, int32_t n_comp,
                            int64_t stride_n, int64_t stride_side, int64_t stride_m, int32_t has_coarse,
                            const uint8_t* valid, const int32_t* idx, int32_t n_rep, double* acc,
                            int64_t acc_rep_stride, void* workspace, int64_t workspace_bytes, cudaStream_t st,
                            const char* who, bool sums_only = false) {
    if (check_basis(basis) != 0) return -1;
    MB_REQUIRE(n >= 0 && n_comp >= 1, "%s: bad n=%lld n_comp=%d", who, (long long)n, n_comp);
    MB_REQUIRE(acc != nullptr && workspace != nullptr, "%s: null acc/workspace", who);
    MB_REQUIRE(n_comp == 1 || valid != nullptr || n_comp <= kThreads,
               "%s: more than %d components need the sample mask (mlmcb200_sample_mask)", who, kThreads);
    if (n == 0) return 0;
    MB_REQUIRE(pairs != nullptr, "%s: null pairs", who);
    const bool gather = idx != nullptr;
    Plan p;
    // few moments + contiguous 16-byte aligned scalar rows + enough samples: HBM-bound -> TMA-ring variant
    const bool fast = n_comp == 1 && valid == nullptr &&
                      (has_coarse ? (stride_n == 2 && stride_side == 1 && (reinterpret_cast<uintptr_t>(pairs) & 15) == 0)
                                  : stride_n >= 1);
    const bool use_ring = fast && !gather && basis->size <= kStreamMaxMoments && basis->kind != MLMCB200_FOURIER &&
                          (stride_n == 2 || (!has_coarse && stride_n == 1)) &&
                          (reinterpret_cast<uintptr_t>(pairs) & 15) == 0 && n >= (int64_t)8 * kThreads * 8;
    // sums only: the scalar FAST kernels have a variant without the sums of squares; everything else computes them anyway
    const bool nosq = sums_only && fast && !gather && !use_ring && basis->kind != MLMCB200_RAW;
    if (plan_moments(basis->kind, basis->size, n_comp, n, has_coarse != 0, use_ring, &p, nosq) != 0) return -1;
    const int64_t K = (int64_t)basis->size * n_comp;
    const int64_t stride = 2 + 2 * K;
    if (gather) {
        p.grid.z = (unsigned)n_rep;
        p.gather = true;
        MB_REQUIRE(workspace_bytes >= mlmcb200_moments_resampled_workspace_bytes(basis->size, n_comp, n_rep),
                   "%s: workspace too small", who);
    } else {
        MB_REQUIRE(workspace_bytes >= (int64_t)p.grid.y * stride * 8, "%s: workspace too small", who);
    }

    MomentsArgs a;
    a.basis = *basis;
    a.pairs = pairs;
    a.valid = valid;
    a.n = n;
    a.stride_n = stride_n;
    a.stride_side = stride_side;
    a.stride_m = stride_m;
    a.n_comp = n_comp;
    a.vec2 = (n_comp == 1 && has_coarse && stride_n == 2 && stride_side == 1 &&
              (reinterpret_cast<uintptr_t>(pairs) & 15) == 0) ? 1 : 0;
    a.partial = static_cast<double*>(workspace);
    a.partial_stride = stride;
    a.idx = idx;
    a.fuse_mask = (n_comp > 1 && valid == nullptr) ? 1 : 0;
    if (a.fuse_mask) {
        p.smem += 1024;                                      // 2 x S x TN flag bytes
        MB_REQUIRE(p.smem + 64 <= 227u * 1024u, "%s: size %d with the in-kernel sample mask exceeds the shared memory",
                   who, basis->size);
    }
    // re-sampling: only the Legendre kernels have a register-blocked gather variant, the other bases take the
    // generic kernel (row indirection at run time)
    p.fast = fast && (!gather || basis->kind == MLMCB200_LEGENDRE);

    int rc = -1;     // > 0: number of partial vectors written (per replicate)
    const bool coarse = has_coarse != 0, is_log = basis->is_log != 0;
    switch (basis->kind) {
        case MLMCB200_RAW: rc = launch_moments_raw(a, p, coarse, st); break;
        case MLMCB200_LEGENDRE:
            rc = coarse ? launch_moments_legendre_coarse(a, p, is_log, st) : launch_moments_legendre_level0(a, p, is_log, st);
            break;
        case MLMCB200_MONOMIAL: rc = launch_moments_monomial(a, p, coarse, is_log, st); break;
        case MLMCB200_FOURIER: rc = launch_moments_fourier(a, p, coarse, is_log, st); break;
    }
    if (rc <= 0) return rc < 0 ? rc : -1;
    return launch_reduce_partials_batched(a.partial, rc, stride, stride, acc, gather ? n_rep : 1, acc_rep_stride, st);
}

}  // namespace

extern "C" int mlmcb200_moments_accumulate(const mlmcb200_basis_t* basis, const double* pairs, int64_t n,
                                           int32_t n_comp, int64_t stride_n, int64_t stride_side,
                                           int64_t stride_m, int32_t has_coarse, const uint8_t* valid,
                                           double* acc, void* workspace, int64_t workspace_bytes, void* stream) {
    return moments_accumulate_impl(basis, pairs, n, n_comp, stride_n, stride_side, stride_m, has_coarse, valid,
                                   nullptr, 1, acc, 0, workspace, workspace_bytes, (cudaStream_t)stream,
                                   "moments_accumulate");
}

extern "C" int mlmcb200_moments_accumulate_sums(const mlmcb200_basis_t* basis, const double* pairs, int64_t n,
                                                int32_t n_comp, int64_t stride_n, int64_t stride_side,
                                                int64_t stride_m, int32_t has_coarse, const uint8_t* valid,
                                                double* acc, void* workspace, int64_t workspace_bytes, void* stream) {
    return moments_accumulate_impl(basis, pairs, n, n_comp, stride_n, stride_side, stride_m, has_coarse, valid,
                                   nullptr, 1, acc, 0, workspace, workspace_bytes, (cudaStream_t)stream,
                                   "moments_accumulate_sums", true);
}

extern "C" int64_t mlmcb200_moments_resampled_workspace_bytes(int32_t size, int32_t n_comp, int32_t n_rep) {
    Plan p;
    if (size < 1 || n_comp < 1 || n_rep < 1) return -1;
    if (plan_moments(MLMCB200_LEGENDRE, size, n_comp, -1, true, false, &p) != 0) return -1;
    // plan grid.y = one wave at the shared-memory bound on CTAs per SM (>= the real occupancy)
    const int64_t ctas = (int64_t)kMaxWavesResampled * p.grid.y + n_rep;
    return ctas * (2 + 2 * (int64_t)size * n_comp) * (int64_t)sizeof(double);
}

extern "C" int mlmcb200_moments_accumulate_resampled(const mlmcb200_basis_t* basis, const double* pairs,
                                                     int64_t n_rows, int32_t n_comp, int64_t stride_n,
                                                     int64_t stride_side, int64_t stride_m, int32_t has_coarse,
                                                     const uint8_t* valid, const int32_t* idx, int64_t n_draws,
                                                     int32_t n_rep, double* acc, int64_t acc_rep_stride,
                                                     void* workspace, int64_t workspace_bytes, void* stream) {
    MB_REQUIRE(idx != nullptr && n_rep >= 1 && n_rep <= 65535, "moments_accumulate_resampled: bad idx / n_rep=%d",
               n_rep);
    MB_REQUIRE(n_rows >= 1 && n_rows <= 0x7fffffffLL, "moments_accumulate_resampled: n_rows=%lld out of range",
               (long long)n_rows);
    MB_REQUIRE(acc_rep_stride >= 2 + 2 * (int64_t)n_comp * (basis ? basis->size : 0),
               "moments_accumulate_resampled: replicate stride too small");
    return moments_accumulate_impl(basis, pairs, n_draws, n_comp, stride_n, stride_side, stride_m, has_coarse, valid,
                                   idx, n_rep, acc, acc_rep_stride, workspace, workspace_bytes,
                                   (cudaStream_t)stream, "moments_accumulate_resampled");
}

extern "C" int mlmcb200_resample_indices(uint64_t seed, uint64_t stream_id, int64_t n_rows, int64_t n_draws,
                                         int32_t n_rep, int32_t rep_offset, int32_t n_blocks, const int64_t* block_cum,
                                         int32_t* idx, void* stream) {
    MB_REQUIRE(n_rows >= 1 && n_rows <= 0x7fffffffLL && n_draws >= 0 && n_rep >= 1 && n_rep <= 65535 && idx != nullptr &&
                   rep_offset >= 0,
               "resample_indices: bad arguments (n_rows=%lld n_draws=%lld n_rep=%d rep_offset=%d)", (long long)n_rows,
               (long long)n_draws, n_rep, rep_offset);
    MB_REQUIRE(n_blocks >= 1 && n_blocks <= n_rows && (n_blocks == 1 || block_cum != nullptr),
               "resample_indices: n_blocks=%d needs the cumulative block counts", n_blocks);
    if (n_draws == 0) return 0;
    const int threads = 256;
    const int64_t pairs = (n_draws + 1) / 2;
    const dim3 grid((unsigned)((pairs + threads - 1) / threads), (unsigned)n_rep);
    resample_indices_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(seed, stream_id, n_rows, n_draws, n_blocks,
                                                                        n_blocks > 1 ? block_cum : nullptr, idx,
                                                                        (uint32_t)rep_offset);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int mlmcb200_sample_mask(const mlmcb200_basis_t* basis, const double* pairs, int64_t n, int32_t n_comp,
                                    int64_t stride_n, int64_t stride_side, int64_t stride_m, int32_t has_coarse,
                                    uint8_t* valid, void* stream) {
    if (check_basis(basis) != 0) return -1;
    MB_REQUIRE(n >= 0 && n_comp >= 1 && valid != nullptr, "sample_mask: bad arguments");
    if (n == 0) return 0;
    const int threads = 256;
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA_OK(cudaMemsetAsync(valid, 1, (size_t)n, st));
    const int64_t total = n * (int64_t)n_comp * (has_coarse ? 2 : 1);
    const int64_t per = (int64_t)n_comp * (has_coarse ? 2 : 1);
    if (stride_m == 1 && (!has_coarse || stride_side == n_comp) && per % 2 == 0 && per >= 64 && per < (1LL << 31) &&
        stride_n % 2 == 0 && (reinterpret_cast<uintptr_t>(pairs) & 15) == 0) {
        const int64_t units = n * ((per / 2 + kMaskSegment - 1) / kMaskSegment);   // one warp per (sample, segment)
        int64_t row_blocks = (units + (threads / 32) - 1) / (threads / 32);
        const int64_t row_cap = (int64_t)sm_count() * 8;
        if (row_blocks > row_cap) row_blocks = row_cap;
        sample_mask_rows_kernel<<<(unsigned)row_blocks, threads, 0, st>>>(*basis, pairs, n, (int)(per / 2), stride_n, valid);
        MB_CUDA_OK(cudaGetLastError());
        return 0;
    }
    int64_t blocks = (total + threads * 8 - 1) / (threads * 8);
    const int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (total < 0xffffffffLL)
        sample_mask_kernel<uint32_t><<<(unsigned)blocks, threads, 0, st>>>(
            *basis, pairs, n, n_comp, stride_n, stride_side, stride_m, has_coarse ? 2 : 1, valid);
    else
        sample_mask_kernel<uint64_t><<<(unsigned)blocks, threads, 0, st>>>(
            *basis, pairs, n, n_comp, stride_n, stride_side, stride_m, has_coarse ? 2 : 1, valid);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int mlmcb200_finalize_levels(const double* acc, int64_t acc_stride, int32_t n_levels, int64_t K,
                                        double* l_means, double* l_vars, double* mean, double* var, void* stream) {
    MB_REQUIRE(acc != nullptr && n_levels >= 1 && K >= 1 && acc_stride >= 2 + 2 * K, "finalize_levels: bad arguments");
    const int threads = 128;
    finalize_levels_kernel<<<(unsigned)((K + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        acc, acc_stride, n_levels, K, l_means, l_vars, mean, var, 0, 0);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int mlmcb200_estimate_moments_levels(const mlmcb200_basis_t* basis, const mlmcb200_level_t* levels,
                                                int32_t n_levels, int32_t n_comp, double* acc, int64_t acc_stride,
                                                double* out, double* host_out, void* workspace,
                                                int64_t workspace_bytes, void* stream) {
    if (check_basis(basis) != 0) return -1;
    MB_REQUIRE(levels != nullptr && n_levels >= 1 && n_comp >= 1 && acc != nullptr && out != nullptr,
               "estimate_moments_levels: bad arguments");
    const int64_t K = (int64_t)n_comp * basis->size;
    MB_REQUIRE(acc_stride >= 2 + 2 * K, "estimate_moments_levels: accumulator stride too small");
    cudaStream_t st = (cudaStream_t)stream;
    MB_CUDA_OK(cudaMemsetAsync(acc, 0, (size_t)n_levels * acc_stride * sizeof(double), st));
    for (int l = 0; l < n_levels; ++l) {
        const mlmcb200_level_t& lv = levels[l];
        if (lv.n == 0) continue;
        const int rc = moments_accumulate_impl(basis, lv.pairs, lv.n, n_comp, lv.stride_n, lv.stride_side, lv.stride_m,
                                               lv.has_coarse, lv.valid, nullptr, 1, acc + (int64_t)l * acc_stride, 0,
                                               workspace, workspace_bytes, st, "estimate_moments_levels");
        if (rc != 0) return rc;
    }
    const int64_t LK = (int64_t)n_levels * K;
    const int threads = 128;
    const int64_t work = K > 2 * n_levels ? K : 2 * n_levels;
    finalize_levels_kernel<<<(unsigned)((work + threads - 1) / threads), threads, 0, st>>>(
        acc, acc_stride, n_levels, K, out, out + LK, out + 2 * LK, out + 2 * LK + K, 0, 0, out + 2 * LK + 2 * K);
    MB_CUDA_OK(cudaGetLastError());
    if (host_out != nullptr) {
        MB_CUDA_OK(cudaMemcpyAsync(host_out, out, (size_t)(2 * LK + 2 * K + 2 * n_levels) * sizeof(double),
                                   cudaMemcpyDeviceToHost, st));
        MB_CUDA_OK(cudaStreamSynchronize(st));
    }
    return 0;
}

extern "C" int mlmcb200_level_sums_transform(const double* acc_in, int64_t in_stride, int32_t n_levels, int32_t K0,
                                             int32_t n_comp, const double* mat_t, int32_t K1, double* acc_out,
                                             int64_t out_stride, void* stream) {
    MB_REQUIRE(acc_in != nullptr && acc_out != nullptr && mat_t != nullptr && n_levels >= 1 && K0 >= 1 && K1 >= 1 &&
                   n_comp >= 1 && (int64_t)n_levels * n_comp <= 65535 &&
                   in_stride >= 2 + 2 * (int64_t)n_comp * K0 && out_stride >= 2 + 2 * (int64_t)n_comp * K1,
               "level_sums_transform: bad arguments");
    const int threads = 128;
    const dim3 grid((unsigned)((K1 + threads - 1) / threads), (unsigned)(n_levels * n_comp));
    level_sums_transform_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(acc_in, in_stride, K0, n_comp, mat_t, K1,
                                                                            acc_out, out_stride);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int mlmcb200_finalize_levels_batched(const double* acc, int64_t acc_stride, int32_t n_levels, int64_t K,
                                                int32_t n_batch, int64_t acc_batch_stride, double* out,
                                                void* stream) {
    MB_REQUIRE(acc != nullptr && out != nullptr && n_levels >= 1 && K >= 1 && acc_stride >= 2 + 2 * K &&
                   n_batch >= 1 && n_batch <= 65535 && acc_batch_stride >= (int64_t)n_levels * acc_stride,
               "finalize_levels_batched: bad arguments");
    const int threads = 128;
    const int64_t LK = (int64_t)n_levels * K;
    const dim3 grid((unsigned)((K + threads - 1) / threads), (unsigned)n_batch);
    finalize_levels_kernel<<<grid, threads, 0, (cudaStream_t)stream>>>(
        acc, acc_stride, n_levels, K, out, out + LK, out + 2 * LK, out + 2 * LK + K, acc_batch_stride, 2 * LK + 2 * K);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}
