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
//
// This header holds the kernel template and its launch helpers; it is included by the moments_k_*.cu translation units,
// each of which instantiates the variants of one basis family.
#pragma once
#include <stdlib.h>
#include "moments_types.cuh"

namespace mlmcb200 {

namespace {

using namespace detail;

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
// vmask: bit 16 + s = sample s exists (in range, active thread); vraw[s] = the caller's sample mask (a.valid) of sample
// s, fetched here -- one tile ahead, with the values -- and left untouched (testing it here would wait for the load).
template <bool COARSE, int S>
__device__ __forceinline__ void load_tile(const MomentsArgs& a, const double* base_f, const int32_t* idx, int64_t n0,
                                          int TN, bool active, double (&xf)[S], double (&xc)[S], unsigned& vmask,
                                          unsigned char (&vraw)[S]) {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    vmask = 0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        int64_t n = n0 + (int64_t)s * TN;
        xf[s] = qnan;
        xc[s] = qnan;
        vraw[s] = 0;
        if (active && n < a.n) {
            vmask |= 0x10000u << s;                                 // bit 16 + s: the sample exists
            if (idx != nullptr) n = __ldg(idx + n);                 // re-sampled row
            if (a.valid != nullptr) vraw[s] = a.valid[n];
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
    unsigned vmask = 0;                                      // generic path: which samples of the tile in flight exist
    unsigned char vraw[S];                                   // ... and the caller's sample mask of each
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
        load_tile<COARSE, S>(a, base_f, idx, (int64_t)blockIdx.y * tile_n + tn, TN, active, xf, xc, vmask, vraw);
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
                const bool in = (vm >> (16 + s)) & 1u, good = vraw[s] != 0;
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
                load_tile<COARSE, S>(a, base_f, idx, (tile + gridDim.y) * tile_n + tn, TN, active, xf, xc, vmask, vraw);
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
                load_tile<COARSE, S>(a, base_f, idx, (tile + gridDim.y) * tile_n + tn, TN, active, xf, xc, vmask, vraw);
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

}  // namespace

}  // namespace mlmcb200
