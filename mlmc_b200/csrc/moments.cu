// Fused generalized-moment evaluation + per-level mean/variance sums (sm_100a).
//
// Replaces the chunk loop of estimate_mean over a `moments` quantity
// (mlmc/quantity/quantity_estimate.py:43-65 with the operation of :105-110, basis from mlmc/moments.py).
// The reference materialises Phi[M, n, 2, R] per chunk, transposes it twice, masks, subtracts and reduces;
// here the basis recurrence runs in registers and only the 2*M*R level sums ever leave the SM.
//
// Work decomposition
//   * one thread owns ONE component m and a strided set of samples; it advances the recurrences of S samples
//     (fine and coarse) in lock step: S*2 independent dependency chains of FP64 work per thread.
//   * per-thread, per-moment running sums (sum d, sum d^2) live in shared memory, [2][R][T] doubles, column
//     `tid` private to the thread (conflict-free 64-bit accesses); a moment's pair is loaded once, updated with
//     the S samples in registers, and stored back: 4 LSU ops per S sample-moments.
//   * block epilogue: tree-reduce the T columns per (m, r) with warp shuffles, un-scale (Legendre), write one
//     partial vector per CTA; `reduce_partials_kernel` adds the partials to the level accumulator in a fixed
//     order (bitwise reproducible, no atomics).
//   * Legendre runs the monic recurrence W_i = t W_{i-1} - e_i W_{i-2} (gen_tables.py) and rescales the sums by
//     g_i / g_i^2 in the epilogue.
//   * FP64 instruction count per level-sample-moment: Legendre 2*(DMUL+DFMA) + DADD + DADD + DFMA = 7
//     (level 0: 4); Monomial 5 (3); Fourier 7 (4).  The kernel is bound by the FP64 pipe, not by HBM, for
//     R >= ~8 (DESIGN.md "Rooflines").
#include <stdlib.h>
#include "common.cuh"

namespace mlmcb200 {

namespace {

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

// ---- per-moment reduction of the S samples held by this thread into its shared-memory column(s) ----
// sum   : pairwise add tree (depth log2 S)
// square: two interleaved FMA chains (depth S/2)
// PAIR (scalar quantity): lanes 2j and 2j+1 share ONE pair of columns -- the even lane keeps the sum of both, the odd
// lane the sum of squares of both (one 64-bit shuffle) -- which halves the shared memory per thread and so doubles
// the number of resident warps for a given number of moments.
// NOSQ (sums only: the caller needs no sums of squares, e.g. the moment sums behind the linearised covariance means):
// one private column per thread, 6 instead of 7.25 FP64 instructions per sample-moment, half the shared memory.
template <bool COARSE, int S, bool PAIR, bool NOSQ = false>
__device__ __forceinline__ void accumulate(double* __restrict__ col, int sq_off, bool odd,
                                           const double (&vf)[S], const double (&vc)[S]) {
    // col = this thread's column entry of the moment (sm + k*T + tid); the squares live sq_off doubles further (private
    // columns) or in the odd lane's entry (lane pairs)
    static_assert(S % 2 == 0, "S must be even");
    double d[S];
#pragma unroll
    for (int s = 0; s < S; ++s) d[s] = COARSE ? vf[s] - vc[s] : vf[s];
    if (NOSQ) {
#pragma unroll
        for (int w = 1; w < S; w <<= 1) {
#pragma unroll
            for (int s = 0; s + w < S; s += 2 * w) d[s] += d[s + w];
        }
        *col += d[0];
        return;
    }
    double qa = d[0] * d[0], qb = d[1] * d[1];
#pragma unroll
    for (int s = 2; s < S; s += 2) {
        qa = fma(d[s], d[s], qa);
        qb = fma(d[s + 1], d[s + 1], qb);
    }
#pragma unroll
    for (int w = 1; w < S; w <<= 1) {
#pragma unroll
        for (int s = 0; s + w < S; s += 2 * w) d[s] += d[s + w];
    }
    const double p1 = d[0], p2 = qa + qb;
    if (PAIR) {
        const double recv = __shfl_xor_sync(0xffffffffu, odd ? p1 : p2, 1);
        *col += (odd ? p2 : p1) + recv;
    } else {
        *col += p1;
        col[sq_off] += p2;
    }
}

// raw (fine, coarse) values of the S samples of one tile owned by this thread; out-of-range samples read NaN.
// vmask: bit s = the caller's sample mask (a.valid) of sample s, fetched here -- one tile ahead, with the values -- so
// that the classification of the tile does not wait for it; bit 16 + s = sample s exists (in range, active thread).
template <bool COARSE, int S>
__device__ __forceinline__ void load_tile(const MomentsArgs& a, const double* base_f, const int32_t* idx, int64_t n0,
                                          int TN, bool active, double (&xf)[S], double (&xc)[S], unsigned& vmask) {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    vmask = 0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        int64_t n = n0 + (int64_t)s * TN;
        xf[s] = qnan;
        xc[s] = qnan;
        if (active && n < a.n) {
            vmask |= 0x10000u << s;                                 // bit 16 + s: the sample exists
            if (idx != nullptr) n = __ldg(idx + n);                 // re-sampled row
            if (a.valid != nullptr) vmask |= (a.valid[n] != 0 ? 1u : 0u) << s;
            if (COARSE && a.vec2) {
                const double2 v = __ldcs(reinterpret_cast<const double2*>(a.pairs) + n);
                xf[s] = v.x;
                xc[s] = v.y;
            } else {
                xf[s] = __ldcs(base_f + n * a.stride_n);
                if (COARSE) xc[s] = __ldcs(base_f + n * a.stride_n + a.stride_side);
            }
        }
    }
}

// FAST: scalar quantity in storage order (pairs contiguous, no external mask): 32-bit tile arithmetic, immediate
// load offsets, no per-sample bounds tests on full tiles.
// STAGES > 0 (FAST only): the tiles of this CTA stream through a ring of STAGES shared-memory buffers filled by TMA bulk
// copies (cp.async.bulk + mbarrier, one elected thread), STAGES - 1 tiles ahead of the compute -- the variant for few
// moments, where the kernel is HBM-bound and the two register-prefetched tiles per warp do not cover the DRAM latency.
// GATHER (FAST, STAGES == 0 only): bootstrap re-sampling, sample i of replicate blockIdx.z is the row idx[z * n + i].
template <int KIND, bool COARSE, bool LOG, int S, bool PAIR, bool FAST, int STAGES, bool GATHER, bool NOSQ>
__global__ void __launch_bounds__(kThreads)
moments_acc_kernel(const MomentsArgs a) {
    extern __shared__ __align__(128) double sm[];
    constexpr int T = kThreads;
    const int tid = threadIdx.x;
    const int R = a.basis.size;
    const int M = a.n_comp;
    static_assert(!(NOSQ && PAIR), "sums-only columns are private");
    const int n_cols = (PAIR || NOSQ) ? R : 2 * R;
    for (int r = 0; r < n_cols; ++r) sm[r * T + tid] = 0.0;         // own column(s) only: no barrier needed

    // thread -> (component, sample lane)
    int m, tn, TN;
    bool active;
    // more than T/2 components: one thread per component (a single sample lane per CTA either way; the epilogue then
    // writes each thread's own columns instead of reducing K outputs over one lane)
    const bool by_comp = 2 * M > T;
    if (by_comp) {
        m = blockIdx.x * T + tid;
        tn = 0;
        TN = 1;
        active = m < M;
    } else {
        TN = T / M;
        tn = tid / M;
        m = tid - tn * M;
        active = tn < TN;
    }
    const int64_t tile_n = (int64_t)S * TN;
    const int64_t n_tiles = (a.n + tile_n - 1) / tile_n;
    const double* const base_f = a.pairs + (int64_t)m * a.stride_m;
    const int32_t* const idx = (GATHER || (!FAST && a.idx != nullptr)) ? a.idx + (int64_t)blockIdx.z * a.n : nullptr;
    const bool count_here = (blockIdx.x == 0) && (by_comp ? tid == 0 : m == 0);
    unsigned cnt_ok = 0, cnt_rm = 0;

    // the moment index is kept out of the vector address arithmetic (column pointer + constant steps) so that the loop
    // counter -- and with it the recurrence coefficients fetched by it -- stays in uniform registers
    double* const col0 = sm + tid;
    const int sq_off = R * T;
    const bool odd_lane = tid & 1;
    // moments are reduced strictly in increasing order: a running column pointer replaces the index arithmetic
#define MB_ACC(K, VF, VC)                                                       \
    {                                                                            \
        accumulate<COARSE, S, PAIR, NOSQ>(col, sq_off, odd_lane, VF, VC);        \
        col += T;                                                                \
    }

    // software pipeline: the raw values of the NEXT tile are in flight while the current tile is reduced
    double xf[S], xc[S];
    unsigned vmask = 0;                                      // generic path: the external sample mask of the tile in flight
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    // FAST addressing: sample (tile, s, tid) = pairs[(tile * S * T + s * T + tid) * stride_n (+ 1 for the coarse half)]
    auto load_fast = [&](int64_t tile) {
        const int64_t first = tile * tile_n + tid;
        const bool full = (tile + 1) * tile_n <= a.n;
        if (GATHER) {
            // two dependent rounds of S independent loads: the row numbers (coalesced), then the rows themselves
            // (rows are shared by all replicates: default caching, no streaming hint)
            int32_t row[S];
#pragma unroll
            for (int s = 0; s < S; ++s) row[s] = (full || first + s * T < a.n) ? __ldg(idx + first + s * T) : -1;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                xf[s] = xc[s] = qnan;
                if (row[s] >= 0) {
                    if (COARSE) {
                        const double2 v = __ldg(reinterpret_cast<const double2*>(a.pairs) + row[s]);
                        xf[s] = v.x;
                        xc[s] = v.y;
                    } else {
                        xf[s] = __ldg(a.pairs + (int64_t)row[s] * a.stride_n);
                    }
                }
            }
        } else if (COARSE) {
            const double2* p = reinterpret_cast<const double2*>(a.pairs) + first;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                double2 v = make_double2(qnan, qnan);
                if (full || first + s * T < a.n) v = __ldcs(p + s * T);
                xf[s] = v.x;
                xc[s] = v.y;
            }
        } else {
            const double* p = a.pairs + first * a.stride_n;
            const int64_t step = (int64_t)T * a.stride_n;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                xf[s] = qnan;
                if (full || first + s * T < a.n) xf[s] = __ldcs(p + s * step);
            }
        }
    };
    // ---- TMA ring (STAGES > 0): stage buffers behind the accumulator columns, one mbarrier per stage ----
    constexpr uint32_t kStageCap = (uint32_t)S * T * 16;                           // bytes reserved per stage
    char* const ring = reinterpret_cast<char*>(sm + (size_t)n_cols * T);
    uint64_t* const bars = reinterpret_cast<uint64_t*>(ring + (size_t)(STAGES > 0 ? STAGES : 1) * kStageCap);
    const uint32_t tile_bytes = (uint32_t)tile_n * 8u * (uint32_t)a.stride_n;      // 8 or 16 bytes per sample
    const int64_t my_tiles = n_tiles > (int64_t)blockIdx.y ? (n_tiles - blockIdx.y + gridDim.y - 1) / gridDim.y : 0;
    auto issue_stage = [&](int64_t k) {                                           // elected thread only
        const int64_t tile = (int64_t)blockIdx.y + k * gridDim.y;
        if ((tile + 1) * tile_n > a.n) return;                                    // ragged last tile: plain loads
        const int st = (int)(k % (STAGES > 0 ? STAGES : 1));
        mbar_expect_tx(&bars[st], tile_bytes);
        bulk_copy_g2s(ring + (size_t)st * kStageCap, reinterpret_cast<const char*>(a.pairs) + tile * (int64_t)tile_bytes,
                      tile_bytes, &bars[st]);
    };
    if (FAST && STAGES > 0) {
        if (tid == 0) {
            for (int st = 0; st < STAGES; ++st) mbar_init(&bars[st], 1);
            mbar_init_fence();
        }
        __syncthreads();
        if (tid == 0)
            for (int64_t k = 0; k < STAGES && k < my_tiles; ++k) issue_stage(k);
    } else if (FAST) {
        load_fast(blockIdx.y);
    } else {
        load_tile<COARSE, S>(a, base_f, idx, (int64_t)blockIdx.y * tile_n + tn, TN, active, xf, xc, vmask);
    }

    int64_t k_tile = 0;
    for (int64_t tile = blockIdx.y; tile < n_tiles; tile += gridDim.y, ++k_tile) {
        double tf[S], tc[S];
        bool ok[S];
        if (FAST) {
            // an out-of-range slot holds NaN: it fails every validity test below, it only must not be counted
            const int64_t first = tile * tile_n + tid;
            const bool full = (tile + 1) * tile_n <= a.n;
            if (STAGES > 0) {
                if (full) {
                    const int st = (int)(k_tile % STAGES);
                    mbar_wait(&bars[st], (uint32_t)((k_tile / STAGES) & 1));
                    const char* stage = ring + (size_t)st * kStageCap;
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        if (COARSE || a.stride_n == 2) {
                            const double2 v = reinterpret_cast<const double2*>(stage)[s * T + tid];
                            xf[s] = v.x;
                            xc[s] = v.y;
                        } else {
                            xf[s] = reinterpret_cast<const double*>(stage)[s * T + tid];
                            xc[s] = 0.0;
                        }
                    }
                } else {
                    load_fast(tile);
                }
                __syncthreads();                                   // every thread has read the stage: refill it
                if (tid == 0 && k_tile + STAGES < my_tiles) issue_stage(k_tile + STAGES);
            }
            // Per-sample preamble.  The common case -- a clipped domain (safe_eval) -- is kept free of branches and of
            // 64-bit index arithmetic on full tiles: affine map (3 FP64 ops per value, as the reference), closed-interval
            // test (NaN fails it), select.  Unclipped bases take the generic validity test (rare NaN replay for Legendre).
            if (KIND != MLMCB200_RAW && a.basis.is_clip) {
                const double lo = a.basis.ref_lo, hi = a.basis.ref_hi;
                const bool always = KIND == MLMCB200_FOURIER && R == 1;          // column 0 is the literal 1
#define MB_CLASSIFY(IN_EXPR)                                                                             \
    _Pragma("unroll") for (int s = 0; s < S; ++s) {                                                      \
        const double vf = LOG ? log(xf[s]) : xf[s];                                                      \
        const double t_f = __dadd_rn(__dmul_rn(__dsub_rn(vf, a.basis.shift), a.basis.scale), lo);        \
        bool good = (t_f >= lo) && (t_f <= hi);                                                          \
        double t_c = 0.0;                                                                                \
        if (COARSE) {                                                                                    \
            const double vc = LOG ? log(xc[s]) : xc[s];                                                  \
            t_c = __dadd_rn(__dmul_rn(__dsub_rn(vc, a.basis.shift), a.basis.scale), lo);                 \
            good = good && (t_c >= lo) && (t_c <= hi);                                                   \
        }                                                                                                \
        const bool in = (IN_EXPR);                                                                       \
        good = (good || always) && in;                                                                   \
        cnt_ok += good ? 1u : 0u;                                                                        \
        cnt_rm += (in && !good) ? 1u : 0u;                                                               \
        tf[s] = good ? t_f : 0.0;                                                                        \
        tc[s] = good ? t_c : 0.0;                                                                        \
        ok[s] = good;                                                                                    \
    }
                if (full) {
                    MB_CLASSIFY(true)
                } else {
                    MB_CLASSIFY(first + s * T < a.n)
                }
#undef MB_CLASSIFY
            } else {
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    tf[s] = KIND == MLMCB200_RAW ? xf[s] : map_to_ref_t<LOG>(a.basis, xf[s]);
                    tc[s] = !COARSE ? 0.0 : (KIND == MLMCB200_RAW ? xc[s] : map_to_ref_t<LOG>(a.basis, xc[s]));
                    bool good = moments_finite(a.basis, tf[s]) && (!COARSE || moments_finite(a.basis, tc[s]));
                    const bool in = full || first + s * T < a.n;
                    good = good && in;
                    cnt_ok += good ? 1u : 0u;
                    cnt_rm += (in && !good) ? 1u : 0u;
                    tf[s] = good ? tf[s] : 0.0;
                    tc[s] = good ? tc[s] : 0.0;
                    ok[s] = good;
                }
            }
            if (STAGES == 0 && tile + gridDim.y < n_tiles) load_fast(tile + gridDim.y);
        } else if (a.valid != nullptr) {
            // External sample mask (wide vector quantities): a kept sample has every component inside the domain, so
            // the affine map needs no test and the classification is bit tests and selects -- no branches, no 64-bit
            // index arithmetic (half of this variant's instructions used to be the generic per-sample tests).
            const unsigned vm = vmask;
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const bool in = (vm >> (16 + s)) & 1u, good = (vm >> s) & 1u;
                double t_f = xf[s], t_c = xc[s];
                if (KIND != MLMCB200_RAW) {
                    const double vf = LOG ? log(t_f) : t_f;
                    t_f = __dadd_rn(__dmul_rn(__dsub_rn(vf, a.basis.shift), a.basis.scale), a.basis.ref_lo);
                    if (COARSE) {
                        const double vc = LOG ? log(t_c) : t_c;
                        t_c = __dadd_rn(__dmul_rn(__dsub_rn(vc, a.basis.shift), a.basis.scale), a.basis.ref_lo);
                    }
                }
                if (count_here) {
                    cnt_ok += good ? 1u : 0u;
                    cnt_rm += (in && !good) ? 1u : 0u;
                }
                tf[s] = good ? t_f : 0.0;
                tc[s] = (COARSE && good) ? t_c : 0.0;
                ok[s] = good;
            }
            if (tile + gridDim.y < n_tiles)
                load_tile<COARSE, S>(a, base_f, idx, (tile + gridDim.y) * tile_n + tn, TN, active, xf, xc, vmask);
        } else {
            const int64_t n0 = tile * tile_n + tn;
            bool own[S];                                   // this component's verdict on the sample, then the sample's
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const int64_t n = n0 + (int64_t)s * TN;
                const bool in = active && n < a.n;
                if (KIND == MLMCB200_RAW) {
                    tf[s] = xf[s];
                    tc[s] = xc[s];
                } else {
                    tf[s] = map_to_ref_t<LOG>(a.basis, xf[s]);
                    tc[s] = COARSE ? map_to_ref_t<LOG>(a.basis, xc[s]) : 0.0;
                }
                own[s] = in && moments_finite(a.basis, tf[s]) && (!COARSE || moments_finite(a.basis, tc[s]));
            }
            if (a.fuse_mask) {
                // mask_nan_samples across the components of a sample: flag per (s, sample lane), double-buffered over
                // tiles so that the reset for tile t+1 cannot overtake a late reader of tile t
                unsigned char* const flag = reinterpret_cast<unsigned char*>(sm + (size_t)n_cols * T) +
                                            (size_t)(k_tile & 1) * (S * TN);
                if (active && m == 0) {
#pragma unroll
                    for (int s = 0; s < S; ++s) flag[s * TN + tn] = 1;
                }
                __syncthreads();
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    if (active && n0 + (int64_t)s * TN < a.n && !own[s]) flag[s * TN + tn] = 0;
                }
                __syncthreads();
#pragma unroll
                for (int s = 0; s < S; ++s) own[s] = active && n0 + (int64_t)s * TN < a.n && flag[s * TN + tn] != 0;
            }
#pragma unroll
            for (int s = 0; s < S; ++s) {
                const bool in = active && n0 + (int64_t)s * TN < a.n;
                const bool good = own[s];
                if (count_here && in) {
                    cnt_ok += good ? 1u : 0u;
                    cnt_rm += good ? 0u : 1u;
                }
                if (!good) {          // a dropped sample contributes exact zeros to every sum
                    tf[s] = 0.0;
                    tc[s] = 0.0;
                }
                ok[s] = good;
            }
            if (tile + gridDim.y < n_tiles)
                load_tile<COARSE, S>(a, base_f, idx, (tile + gridDim.y) * tile_n + tn, TN, active, xf, xc, vmask);
        }

        double* col = col0;                                      // column entry of the next moment to reduce
        if (KIND == MLMCB200_RAW) {
            MB_ACC(0, tf, tc);
        } else if (KIND == MLMCB200_LEGENDRE) {
            // Monic recurrence W_k = t W_{k-1} - e_k W_{k-2} (P_k = g_k W_k is applied in the epilogue), kept as
            //     W_k = fma(-e_k, W_{k-2}, Z),  Z = t * W_{k-1}          (Z carried in a register)
            // because a DFMA with three distinct register operands issues at 2/3 rate on sm_100 (24.7 vs 36.5 TFLOP/s,
            // tools/fp64_probe.py): here both instructions read two registers (e_k sits in a uniform register).
            // Even k live in (fb, cb), odd k in (fa, ca).  The reduction of moment k is issued AFTER the recurrence of
            // moment k+1 so that its dependent tail overlaps independent FP64 work.
            double fa[S], fb[S], ca[S], cb[S], zf[S], zc[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                fb[s] = cb[s] = ok[s] ? 1.0 : 0.0;
                fa[s] = tf[s];
                ca[s] = tc[s];
                zf[s] = tf[s] * tf[s];
                zc[s] = tc[s] * tc[s];
            }
            MB_ACC(0, fb, cb);
#define MB_REC(DST_F, DST_C, E)                                                  \
    _Pragma("unroll") for (int s = 0; s < S; ++s) {                              \
        DST_F[s] = fma(-(E), DST_F[s], zf[s]);                                   \
        if (COARSE) DST_C[s] = fma(-(E), DST_C[s], zc[s]);                       \
    }                                                                            \
    _Pragma("unroll") for (int s = 0; s < S; ++s) {                              \
        zf[s] = tf[s] * DST_F[s];                                                \
        if (COARSE) zc[s] = tc[s] * DST_C[s];                                    \
    }
            int i = 2;                                                            // next moment to generate
            for (; i + 3 < R; i += 4) {
                // e_k are read straight from constant memory with the (uniform) loop index so that they stay in
                // UNIFORM registers: as a vector-register operand they would be the third register read of the DFMA
                const double e0 = kLegCoef[i], e1 = kLegCoef[i + 1], e2 = kLegCoef[i + 2], e3 = kLegCoef[i + 3];
                MB_REC(fb, cb, e0)
                MB_ACC(i - 1, fa, ca);
                MB_REC(fa, ca, e1)
                MB_ACC(i, fb, cb);
                MB_REC(fb, cb, e2)
                MB_ACC(i + 1, fa, ca);
                MB_REC(fa, ca, e3)
                MB_ACC(i + 2, fb, cb);
            }
            for (; i < R; ++i) {                                                 // remainder, one moment at a time
                const double e = kLegCoef[i];
                if ((i & 1) == 0) {
                    MB_REC(fb, cb, e)
                    MB_ACC(i - 1, fa, ca);
                } else {
                    MB_REC(fa, ca, e)
                    MB_ACC(i - 1, fb, cb);
                }
            }
            if (R > 1) {                                                          // the last moment is still pending
                if ((R - 1) & 1) MB_ACC(R - 1, fa, ca)
                else MB_ACC(R - 1, fb, cb)
            }
#undef MB_REC
        } else if (KIND == MLMCB200_MONOMIAL) {
            // t^k = t^{k-1} t on two ping-pong sets, reduction of moment k issued after the product for k+1
            double fa[S], fb[S], ca[S], cb[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                fb[s] = cb[s] = ok[s] ? 1.0 : 0.0;
                fa[s] = tf[s];
                ca[s] = tc[s];
            }
            MB_ACC(0, fb, cb);
            for (int i = 2; i < R; ++i) {
                if ((i & 1) == 0) {
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        fb[s] = fa[s] * tf[s];
                        if (COARSE) cb[s] = ca[s] * tc[s];
                    }
                    MB_ACC(i - 1, fa, ca);
                } else {
#pragma unroll
                    for (int s = 0; s < S; ++s) {
                        fa[s] = fb[s] * tf[s];
                        if (COARSE) ca[s] = cb[s] * tc[s];
                    }
                    MB_ACC(i - 1, fb, cb);
                }
            }
            if (R > 1) {
                if ((R - 1) & 1) MB_ACC(R - 1, fa, ca)
                else MB_ACC(R - 1, fb, cb)
            }
        } else {  // FOURIER: columns 1, cos t, sin t, cos 2t, sin 2t, ... by exact-angle rotation
            double cf1[S], sf1[S], cc1[S], sc1[S], cfk[S], sfk[S], cck[S], sck[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                cfk[s] = cck[s] = ok[s] ? 1.0 : 0.0;
                cf1[s] = sf1[s] = cc1[s] = sc1[s] = 0.0;
                if (ok[s] && R > 1) {
                    sincos_bounded(tf[s], &sf1[s], &cf1[s]);
                    if (COARSE) sincos_bounded(tc[s], &sc1[s], &cc1[s]);
                }
            }
            MB_ACC(0, cfk, cck);
#pragma unroll
            for (int s = 0; s < S; ++s) {
                cfk[s] = cf1[s];
                sfk[s] = sf1[s];
                cck[s] = cc1[s];
                sck[s] = sc1[s];
            }
            for (int i = 1; i < R; i += 2) {
                MB_ACC(i, cfk, cck);
                if (i + 1 < R) MB_ACC(i + 1, sfk, sck)
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    const double nc = fma(cfk[s], cf1[s], -(sfk[s] * sf1[s]));
                    sfk[s] = fma(sfk[s], cf1[s], cfk[s] * sf1[s]);
                    cfk[s] = nc;
                    if (COARSE) {
                        const double mc = fma(cck[s], cc1[s], -(sck[s] * sc1[s]));
                        sck[s] = fma(sck[s], cc1[s], cck[s] * sc1[s]);
                        cck[s] = mc;
                    }
                }
            }
        }
    }
#undef MB_ACC

    // ---------------- block epilogue ----------------
    __syncthreads();
    double* const out = a.partial + ((int64_t)blockIdx.z * gridDim.y + blockIdx.y) * a.partial_stride;
    const int64_t K = (int64_t)M * R;
    if (by_comp) {
        if (active) {
            for (int r = 0; r < R; ++r) {
                const double al = (KIND == MLMCB200_LEGENDRE) ? kLegAlpha[r] : 1.0;
                out[2 + (int64_t)m * R + r] = sm[r * T + tid] * al;
                out[2 + K + (int64_t)m * R + r] = NOSQ ? 0.0 : sm[(R + r) * T + tid] * (al * al);
            }
        }
    } else if (!PAIR && TN <= 16) {
        // few sample lanes per component: one THREAD per output sums its lanes serially; outputs are walked components
        // fastest, so consecutive threads read consecutive shared-memory columns (a warp per output would spend
        // two shuffle trees on 2..16 values and cost as much as the main loop at M = 64)
        for (int o = tid; o < (int)K; o += T) {
            const int r = o / M, mm = o - r * M;
            double s1 = 0.0, s2 = 0.0;
            for (int j = 0; j < TN; ++j) {
                s1 += sm[r * T + j * M + mm];
                if (!NOSQ) s2 += sm[(R + r) * T + j * M + mm];
            }
            const double al = (KIND == MLMCB200_LEGENDRE) ? kLegAlpha[r] : 1.0;
            out[2 + (int64_t)mm * R + r] = s1 * al;
            out[2 + K + (int64_t)mm * R + r] = s2 * (al * al);
        }
    } else {
        const int warp = tid >> 5, lane = tid & 31, n_warps = T >> 5;
        for (int k = warp; k < (int)K; k += n_warps) {
            const int mm = k / R, r = k - mm * R;
            double s1 = 0.0, s2 = 0.0;
            if (PAIR) {                         // M == 1: even columns hold sums, odd columns sums of squares
                for (int j = 2 * lane; j < T; j += 64) {
                    s1 += sm[r * T + j];
                    s2 += sm[r * T + j + 1];
                }
            } else {
                for (int j = lane; j < TN; j += 32) {
                    s1 += sm[r * T + j * M + mm];
                    if (!NOSQ) s2 += sm[(R + r) * T + j * M + mm];
                }
            }
            s1 = warp_sum(s1);
            s2 = warp_sum(s2);
            if (lane == 0) {
                const double al = (KIND == MLMCB200_LEGENDRE) ? kLegAlpha[r] : 1.0;
                out[2 + k] = s1 * al;
                out[2 + K + k] = s2 * (al * al);
            }
        }
    }
    // sample counts: integer-exact
    if (blockIdx.x == 0) {
        __shared__ unsigned cnt_sm[2];
        if (tid == 0) cnt_sm[0] = cnt_sm[1] = 0;
        __syncthreads();
        if (cnt_ok) atomicAdd(&cnt_sm[0], cnt_ok);
        if (cnt_rm) atomicAdd(&cnt_sm[1], cnt_rm);
        __syncthreads();
        if (tid == 0) {
            out[0] = (double)cnt_sm[0];
            out[1] = (double)cnt_sm[1];
        }
    }
}

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

// ---- row numbers of the bootstrap replicates (counter-based Philox4x32-10: reproducible, no state) ----
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        c[0] = hi1 ^ c[1] ^ k0;
        c[1] = lo1;
        c[2] = hi0 ^ c[3] ^ k1;
        c[3] = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}

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

// Sample partitions (grid.y) of a launch with `columns` CTA columns (component blocks x bootstrap replicates): chosen so
// that columns x partitions fills whole resident waves of `slots` CTAs -- a 1.07-wave grid runs at half speed.
constexpr unsigned kMaxWavesResampled = 16;
constexpr unsigned kMaxWavesVector = 2;
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

template <int KIND, bool COARSE, bool LOG, int S, bool PAIR, bool FAST, int STAGES = 0, bool GATHER = false,
          bool NOSQ = false>
int launch_moments(const MomentsArgs& a, Plan p, cudaStream_t st) {
    auto kern = moments_acc_kernel<KIND, COARSE, LOG, S, PAIR, FAST, STAGES, GATHER, NOSQ>;
    if (STAGES > 0) p.smem += (size_t)STAGES * ((size_t)S * kThreads * 16 + 8);
    // resident CTAs per SM for this variant and shared-memory size (registers may bind before shared memory);
    // the sample-partition dimension of the grid is sized to exactly one resident wave
    // (both are per-device properties of the function: the cache is keyed by the current device as well)
    static thread_local size_t cached_smem = 0;
    static thread_local int cached_ctas = 0, cached_dev = -1;
    int cur_dev = 0;
    MB_CUDA_OK(cudaGetDevice(&cur_dev));
    if (cached_smem != p.smem || cached_ctas == 0 || cached_dev != cur_dev) {
        MB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
        int ctas = 0;
        MB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, kThreads, p.smem));
        cached_ctas = ctas > 0 ? ctas : 1;
        cached_smem = p.smem;
        cached_dev = cur_dev;
    }
    const unsigned slots = (unsigned)(sm_count() * cached_ctas);
    const int TN = a.n_comp >= kThreads ? 1 : kThreads / a.n_comp;
    const int64_t tiles = (a.n + (int64_t)S * TN - 1) / ((int64_t)S * TN);
    if (p.grid.z > 1) {
        p.grid.y = choose_partitions(slots, p.grid.x * p.grid.z, tiles, kMaxWavesResampled);
    } else if (p.grid.x > 1) {
        const unsigned gy = choose_partitions(slots, p.grid.x, tiles, kMaxWavesVector);
        if (gy < p.grid.y) p.grid.y = gy;
    } else if (p.grid.y > slots) {
        p.grid.y = slots;
    }
    kern<<<p.grid, kThreads, p.smem, st>>>(a);
    MB_CUDA_OK(cudaGetLastError());
    return (int)p.grid.y;
}

template <int KIND, bool COARSE, bool LOG, int S>
int launch_moments_pair(const MomentsArgs& a, const Plan& p, cudaStream_t st) {
    // `if constexpr` keeps the number of kernel instantiations down (each costs ~1 s of build time)
    if constexpr (KIND != MLMCB200_FOURIER && S == 8) {
        if (p.fast && p.stream && !p.pair) return launch_moments<KIND, COARSE, LOG, S, false, true, kStages>(a, p, st);
    }
    if constexpr (KIND == MLMCB200_LEGENDRE) {
        if (p.fast && p.gather)
            return p.pair ? launch_moments<KIND, COARSE, LOG, S, true, true, 0, true>(a, p, st)
                          : launch_moments<KIND, COARSE, LOG, S, false, true, 0, true>(a, p, st);
    }
    if constexpr (KIND != MLMCB200_RAW) {
        if (p.fast && p.nosq) return launch_moments<KIND, COARSE, LOG, S, false, true, 0, false, true>(a, p, st);
        if (p.fast)
            return p.pair ? launch_moments<KIND, COARSE, LOG, S, true, true>(a, p, st)
                          : launch_moments<KIND, COARSE, LOG, S, false, true>(a, p, st);
    }
    return p.pair ? launch_moments<KIND, COARSE, LOG, S, true, false>(a, p, st)
                  : launch_moments<KIND, COARSE, LOG, S, false, false>(a, p, st);
}

template <int KIND, bool COARSE, bool LOG>
int launch_moments_s(const MomentsArgs& a, const Plan& p, cudaStream_t st) {
    if constexpr (KIND == MLMCB200_FOURIER) {
        return launch_moments_pair<KIND, COARSE, LOG, 4>(a, p, st);
    } else {
        if constexpr (!COARSE && KIND != MLMCB200_RAW) {
            if (p.S == 16) return launch_moments_pair<KIND, false, LOG, 16>(a, p, st);
        }
        return launch_moments_pair<KIND, COARSE, LOG, 8>(a, p, st);
    }
}

template <int KIND>
int launch_moments_kind(const MomentsArgs& a, const Plan& p, bool coarse, bool is_log, cudaStream_t st) {
    if constexpr (KIND == MLMCB200_RAW) {
        return coarse ? launch_moments_s<KIND, true, false>(a, p, st) : launch_moments_s<KIND, false, false>(a, p, st);
    }
    if (is_log)
        return coarse ? launch_moments_s<KIND, true, true>(a, p, st) : launch_moments_s<KIND, false, true>(a, p, st);
    return coarse ? launch_moments_s<KIND, true, false>(a, p, st) : launch_moments_s<KIND, false, false>(a, p, st);
}

}  // namespace

}  // namespace mlmcb200

using namespace mlmcb200;

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
        case MLMCB200_RAW: rc = launch_moments_kind<MLMCB200_RAW>(a, p, coarse, is_log, st); break;
        case MLMCB200_LEGENDRE: rc = launch_moments_kind<MLMCB200_LEGENDRE>(a, p, coarse, is_log, st); break;
        case MLMCB200_MONOMIAL: rc = launch_moments_kind<MLMCB200_MONOMIAL>(a, p, coarse, is_log, st); break;
        case MLMCB200_FOURIER: rc = launch_moments_kind<MLMCB200_FOURIER>(a, p, coarse, is_log, st); break;
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
