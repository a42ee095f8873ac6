// All bootstrap replicates of a level in ONE pass over its samples (sm_100a): weighted level sums on FP64 tensor tiles.
//
// Replaces the replicate loop of Estimate.est_bootstrap (mlmc/estimator.py:171-218) over Quantity.pick_samples
// (mlmc/quantity/quantity.py:307-322: rows drawn WITH replacement) + estimate_mean of the moments
// (mlmc/quantity/quantity_estimate.py:43-65).  A replicate that draws row i  w_bi  times contributes
//     sum_i w_bi d_r(i)      and      sum_i w_bi d_r(i)^2            (d_r = phi_r(fine) - phi_r(coarse))
// to its level sums, i.e. all replicates together are the dense product  W^T [ok | rm | D | D.D]  with W [n, B] the
// multiplicities (mlmcb200_resample_counts: the SAME Philox draws as mlmcb200_resample_indices, histogrammed in shared
// memory) and D [n, R] the moment differences -- SURVEY.md 8f rank 1: "all replicates can be fused into one pass with
// per-replicate multinomial weights".  The basis recurrence of a sample runs ONCE for all replicates (the gather kernel
// re-evaluates it per replicate: 7.25 FP64 instructions per replicate-sample-moment against 2 flop on the tensor pipe
// here) and the rows are read once, in storage order, instead of B times through L2 in random order.
//
// Tile: 192 samples.  X [192][108] doubles in shared memory (columns: ok, rm, d_0.., d_0^2.., zero padding; row stride
// 4 mod 8 and rows skewed by (s / 4) % 4 as in gram.cu: conflict-free fragment loads and producer stores; a sample is
// produced by two threads, one per moment parity, each running the fine and the coarse chain side by side), the
// multiplicities of up to 128 replicates as bytes [replicate][208].  m8n8k4 fragments (lane l): A[m = replicate l/4]
// [k = sample l%4] = count byte, B[k = sample l%4][n = column l/4] = X element, C[replicate l/4][columns 2(l%4), +1].
// A warp owns 8 replicates x all column blocks (<= 26 accumulator registers); when the number of replicate blocks is
// 1 or 2 (mod 4) the last one / two are cut into column ranges over 4 warps so that the four SM sub-partitions (each
// with its own FP64 tensor pipe) carry the same number of DMMAs (100 replicates x 50 moments: 42-43 per k-step each).
#include <stdlib.h>
#include "common.cuh"
#include "philox.cuh"

namespace mlmcb200 {

namespace {

constexpr int kBsThreads = 512;
constexpr int kBsTile = 192;            // samples per tile (12 producing warps; X tile 166 kB)
constexpr int kBsLD = 108;              // row stride of the X tile (doubles)
constexpr int kBsMaxColBlocks = 13;     // 104 columns >= 2 + 2 R  ->  R <= 51
constexpr int kBsMaxRepBlocks = 16;     // 128 replicates per CTA (grid.y runs over groups of replicates)
constexpr int kBsPitch = kBsTile + 16;   // bytes per replicate row of the count tile (16-byte aligned; 208 = 52 words:
                                        // the 8 rows of a fragment start 20 banks apart -> 8 disjoint 4-bank groups)
constexpr int kBsPieces = kBsTile / 16; // 16-byte pieces per replicate row
constexpr int kBsPiecesPerThread = kBsMaxRepBlocks * 8 * kBsPieces / kBsThreads;
static_assert(kBsMaxRepBlocks * 8 * kBsPieces % kBsThreads == 0, "count tile pieces per thread");
constexpr int kBsOutPitch = 8 * kBsMaxColBlocks;
constexpr int kCountCapRows = 196608;   // rows per row block of mlmcb200_resample_counts: byte counters in 192 kB

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

struct WeightedArgs {
    mlmcb200_basis_t basis;
    const double* pairs;
    int64_t n, stride_n;
    const uint8_t* counts;      // [n_rep][counts_stride]
    int64_t counts_stride;
    int32_t n_rep;              // all replicates of the call
    int32_t reps_per_group;     // multiple of 8, <= 128; group g = replicates [g * reps_per_group, ...)
    double* partial;            // [column group][replicate group][gridDim.x][128][104]
};

// The contraction of one tile for a warp that owns NC column blocks of one replicate block.
template <int NC>
__device__ __forceinline__ void contract_tile(double (&acc)[kBsMaxColBlocks][2], const double* __restrict__ xb,
                                              const uint8_t* __restrict__ wrow, int shift) {
#pragma unroll 1
    for (int k16 = 0; k16 < kBsTile / 16; ++k16) {
        const uint4 w = *reinterpret_cast<const uint4*>(wrow + 16 * k16);      // counts of 16 samples of this replicate
        const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {                                          // k-step 4 * k16 + q: row skew q
            const double a = (double)((ww[q] >> shift) & 0xffu);
            const double* xk = xb + (size_t)(16 * k16 + 4 * q) * kBsLD + q;
            double b[NC];
#pragma unroll
            for (int j = 0; j < NC; ++j) b[j] = xk[8 * j];
#pragma unroll
            for (int j = 0; j < NC; ++j) dmma(acc[j][0], acc[j][1], a, b[j]);
        }
    }
}

// Two steps of the monic Legendre recurrence at once, W_j = (t^2 - A_j) W_{j-2} - B_j W_{j-4} (gen_tables.py; gram.cu
// uses the same split): the even and the odd moments of a sample are independent chains, so TWO threads produce a sample
// (2 x 192 threads = 12 warps advance the tile instead of 6), each running the fine and the coarse chain side by side.
static __constant__ double kLegA2[MLMCB200_MAX_MOMENTS] = MLMCB200_LEG_A2_INIT;
static __constant__ double kLegB2[MLMCB200_MAX_MOMENTS] = MLMCB200_LEG_B2_INIT;

// d_j = W_j(fine) - W_j(coarse) and d_j^2 for j = PARITY, PARITY + 2, ... < R into the columns 2 + j and 2 + R + j of the
// sample's row.  PARITY is a template parameter and the caller branches warp-uniformly, so j and the coefficients stay in
// uniform registers.  tf = tc = 0 and good = false for a dropped sample: it starts from W = 0 and writes exact zeros.
// MULTI: more than 104 columns -- the CTA holds the column group [c0, c0 + 104) only (blockIdx.z), stores outside it
// are dropped (the recurrence runs in full either way: the squares need every d_j).
template <bool COARSE, int PARITY, bool MULTI>
__device__ __forceinline__ void produce_moments(int R, bool mono, int c0, double tf, double tc, bool good, double* row) {
    if (PARITY >= R) return;
    auto put = [&](int col, double v) {
        if (!MULTI) row[col] = v;
        else if ((unsigned)(col - c0) < (unsigned)kBsOutPitch) row[col - c0] = v;
    };
    const double uf = tf * tf, uc = tc * tc;
    double f0 = 0.0, c0v = 0.0;                                                  // W_{j-4}
    double f1 = good ? (PARITY ? tf : 1.0) : 0.0, c1 = good ? (PARITY ? tc : 1.0) : 0.0;   // W_{j-2}
    {
        const double d = COARSE ? f1 - c1 : f1;
        put(2 + PARITY, d);
        put(2 + R + PARITY, d * d);
    }
    for (int j = PARITY + 2; j < R; j += 2) {
        const double a2 = mono ? 0.0 : kLegA2[j], b2 = mono ? 0.0 : kLegB2[j];
        const double qf = fma(uf - a2, f1, -(b2 * f0));
        const double qc = COARSE ? fma(uc - a2, c1, -(b2 * c0v)) : 0.0;
        f0 = f1;
        f1 = qf;
        c0v = c1;
        c1 = qc;
        const double d = COARSE ? qf - qc : qf;
        put(2 + j, d);
        put(2 + R + j, d * d);
    }
}

// Fourier basis (columns 1, cos t, sin t, cos 2t, sin 2t, ...: mlmc/moments.py Fourier.eval_all): the two threads of a
// sample take the odd and the even harmonics, each rotating its (cos, sin) pair by the exact double angle 2t per step.
template <bool COARSE, int PARITY, bool MULTI>
__device__ __forceinline__ void produce_fourier(int R, int c0, double tf, double tc, bool good, double* row) {
    auto put = [&](int col, double v) {
        if (!MULTI) row[col] = v;
        else if ((unsigned)(col - c0) < (unsigned)kBsOutPitch) row[col - c0] = v;
    };
    if (PARITY == 0) {                                   // column 0 is the literal 1: its difference vanishes
        const double d = COARSE ? 0.0 : (good ? 1.0 : 0.0);
        put(2, d);
        put(2 + R, d * d);
    }
    if (R < 2) return;
    double sf1 = 0.0, cf1 = 0.0, sc1 = 0.0, cc1 = 0.0;   // a dropped sample keeps zeros and so writes zeros
    if (good) {
        sincos_bounded(tf, &sf1, &cf1);
        if (COARSE) sincos_bounded(tc, &sc1, &cc1);
    }
    const double cf2 = fma(cf1, cf1, -(sf1 * sf1)), sf2 = 2.0 * sf1 * cf1;
    const double cc2 = fma(cc1, cc1, -(sc1 * sc1)), sc2 = 2.0 * sc1 * cc1;
    double cf = PARITY ? cf2 : cf1, sf = PARITY ? sf2 : sf1, cc = PARITY ? cc2 : cc1, sc = PARITY ? sc2 : sc1;
    for (int k = 1 + PARITY; 2 * k - 1 < R; k += 2) {     // harmonic k: columns 2k - 1 (cos), 2k (sin)
        const double dcos = COARSE ? cf - cc : cf;
        put(2 + 2 * k - 1, dcos);
        put(2 + R + 2 * k - 1, dcos * dcos);
        if (2 * k < R) {
            const double dsin = COARSE ? sf - sc : sf;
            put(2 + 2 * k, dsin);
            put(2 + R + 2 * k, dsin * dsin);
        }
        const double nf = fma(cf, cf2, -(sf * sf2));
        sf = fma(sf, cf2, cf * sf2);
        cf = nf;
        if (COARSE) {
            const double nc = fma(cc, cc2, -(sc * sc2));
            sc = fma(sc, cc2, cc * sc2);
            cc = nc;
        }
    }
}

template <bool COARSE, bool MULTI>
__global__ void __launch_bounds__(kBsThreads, 1)
weighted_moments_kernel(const WeightedArgs a) {
    extern __shared__ __align__(16) double smem_d[];
    double* const X = smem_d;                                                   // [kBsTile][108] + skew
    uint8_t* const Wt = reinterpret_cast<uint8_t*>(smem_d + kBsTile * kBsLD + 8);  // [128][kBsPitch]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int R = a.basis.size;
    const int c0 = MULTI ? (int)blockIdx.z * kBsOutPitch : 0;                    // first column of this CTA's group
    const int ncb = (min(kBsOutPitch, 2 + 2 * R - c0) + 7) >> 3;
    const bool mono = a.basis.kind == MLMCB200_MONOMIAL, fourier = a.basis.kind == MLMCB200_FOURIER;
    const int rep0 = blockIdx.y * a.reps_per_group;
    const int n_rep = min(a.reps_per_group, a.n_rep - rep0);                     // replicates of this group
    const int nbb = (n_rep + 7) >> 3;

    // warp -> (replicate block, column range)
    const int rem = nbb & 3;
    const int n_full = rem == 1 ? nbb - 1 : rem == 2 ? nbb - 2 : nbb;
    int brow = -1, c_lo = 0, nc = 0;
    if (warp < n_full) {
        brow = warp;
        nc = ncb;
    } else if (rem == 1 && warp < n_full + 4) {
        const int piece = warp - n_full;
        brow = nbb - 1;
        c_lo = piece * ncb / 4;
        nc = (piece + 1) * ncb / 4 - c_lo;
    } else if (rem == 2 && warp < n_full + 4) {
        const int piece = warp - n_full;
        brow = nbb - 2 + (piece >> 1);
        c_lo = (piece & 1) * ncb / 2;
        nc = ((piece & 1) + 1) * ncb / 2 - c_lo;
    }

    // zero the tile once: the padding columns are never written again, count rows of absent replicates stay 0
    for (int i = tid; i < kBsTile * kBsLD + 8; i += kBsThreads) X[i] = 0.0;
    for (int i = tid; i < kBsMaxRepBlocks * 8 * kBsPitch / 4; i += kBsThreads) reinterpret_cast<uint32_t*>(Wt)[i] = 0u;

    double acc[kBsMaxColBlocks][2];
#pragma unroll
    for (int j = 0; j < kBsMaxColBlocks; ++j) acc[j][0] = acc[j][1] = 0.0;

    const int64_t n_tiles = (a.n + kBsTile - 1) / kBsTile;
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    // producers: threads s and kBsTile + s produce the even / the odd moments of sample s (parity is warp-uniform)
    const int parity = tid >= kBsTile ? 1 : 0;
    const int ps = tid - parity * kBsTile;
    const bool producer = tid < 2 * kBsTile;
    auto fetch_value = [&](int64_t tile) {
        const int64_t s = tile * kBsTile + ps;
        if (!producer || s >= a.n) return make_double2(qnan, qnan);
        if (COARSE) return __ldcs(reinterpret_cast<const double2*>(a.pairs) + s);
        return make_double2(__ldcs(a.pairs + s * a.stride_n), 0.0);
    };
    // count tile: 128 replicates x kBsPieces pieces of 16 samples, kBsPiecesPerThread per thread
    auto fetch_counts = [&](int64_t tile, int piece) {
        const int i = tid + piece * kBsThreads, r = i / kBsPieces, o = (i - r * kBsPieces) * 16;
        const int64_t s0 = tile * kBsTile + o;
        if (r >= n_rep || s0 + 16 > a.counts_stride) return make_uint4(0u, 0u, 0u, 0u);
        return __ldcs(reinterpret_cast<const uint4*>(a.counts + (int64_t)(rep0 + r) * a.counts_stride + s0));
    };
    double2 v_next = fetch_value(blockIdx.x);
    uint4 w_next[kBsPiecesPerThread];
#pragma unroll
    for (int piece = 0; piece < kBsPiecesPerThread; ++piece) w_next[piece] = fetch_counts(blockIdx.x, piece);
    __syncthreads();

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        // ---------------- produce ----------------
#pragma unroll
        for (int piece = 0; piece < kBsPiecesPerThread; ++piece) {
            const int i = tid + piece * kBsThreads, r = i / kBsPieces, o = (i - r * kBsPieces) * 16;
            *reinterpret_cast<uint4*>(Wt + r * kBsPitch + o) = w_next[piece];
        }
        if (producer) {
            const int64_t s = tile * kBsTile + ps;
            const bool in = s < a.n;
            double tf = a.basis.is_log ? log(v_next.x) : v_next.x;
            tf = __dadd_rn(__dmul_rn(__dsub_rn(tf, a.basis.shift), a.basis.scale), a.basis.ref_lo);
            bool good = a.basis.is_clip ? (tf >= a.basis.ref_lo && tf <= a.basis.ref_hi) : moments_finite(a.basis, tf);
            double tc = 0.0;
            if (COARSE) {
                tc = a.basis.is_log ? log(v_next.y) : v_next.y;
                tc = __dadd_rn(__dmul_rn(__dsub_rn(tc, a.basis.shift), a.basis.scale), a.basis.ref_lo);
                const bool good_c = a.basis.is_clip ? (tc >= a.basis.ref_lo && tc <= a.basis.ref_hi)
                                                    : moments_finite(a.basis, tc);
                good = good && good_c;
            }
            good = (good || (fourier && R == 1)) && in;   // a Fourier basis of one function is the literal 1: nothing to drop
            tf = good ? tf : 0.0;
            tc = good ? tc : 0.0;
            double* const row = X + (size_t)ps * kBsLD + ((ps >> 2) & 3);
            if (parity == 0 && c0 == 0) {
                row[0] = good ? 1.0 : 0.0;
                row[1] = (in && !good) ? 1.0 : 0.0;
            }
            if (fourier) {                                // uniform
                if (parity == 0) produce_fourier<COARSE, 0, MULTI>(R, c0, tf, tc, good, row);
                else produce_fourier<COARSE, 1, MULTI>(R, c0, tf, tc, good, row);
            } else {
                if (parity == 0) produce_moments<COARSE, 0, MULTI>(R, mono, c0, tf, tc, good, row);
                else produce_moments<COARSE, 1, MULTI>(R, mono, c0, tf, tc, good, row);
            }
        }
        __syncthreads();
        // ---------------- prefetch the next tile, contract this one ----------------
        const int64_t nxt = tile + gridDim.x;
        if (nxt < n_tiles) {
            v_next = fetch_value(nxt);
#pragma unroll
            for (int piece = 0; piece < kBsPiecesPerThread; ++piece) w_next[piece] = fetch_counts(nxt, piece);
        }
        if (brow >= 0) {
            const double* xb = X + (size_t)(lane & 3) * kBsLD + (lane >> 2) + 8 * c_lo;
            const uint8_t* wrow = Wt + (brow * 8 + (lane >> 2)) * kBsPitch;
            const int shift = 8 * (lane & 3);
            switch (nc) {
#define MB_CASE(N) case N: contract_tile<N>(acc, xb, wrow, shift); break;
                MB_CASE(1) MB_CASE(2) MB_CASE(3) MB_CASE(4) MB_CASE(5) MB_CASE(6) MB_CASE(7)
                MB_CASE(8) MB_CASE(9) MB_CASE(10) MB_CASE(11) MB_CASE(12) MB_CASE(13)
#undef MB_CASE
                default: break;
            }
        }
        __syncthreads();
    }

    // per-CTA partial [128][104]: C fragment (lane l): row l/4, columns 2 (l%4), 2 (l%4) + 1
    if (brow >= 0) {
        double* out = a.partial + (((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) *
                                      (kBsMaxRepBlocks * 8 * kBsOutPitch) +
                      (size_t)(brow * 8 + (lane >> 2)) * kBsOutPitch + 8 * c_lo + 2 * (lane & 3);
#pragma unroll
        for (int j = 0; j < kBsMaxColBlocks; ++j) {
            if (j < nc) *reinterpret_cast<double2*>(out + 8 * j) = make_double2(acc[j][0], acc[j][1]);
        }
    }
}

// acc[b][...] += (sum over the CTAs' partials, in CTA order), Legendre re-scaled to P_k = alpha_k W_k; one thread per
// (b, column)
__global__ void weighted_finish_kernel(const double* __restrict__ partial, int n_partials, int n_rep, int reps_per_group,
                                       int n_rep_groups, int R, int unscaled, double* __restrict__ acc,
                                       int64_t acc_rep_stride) {
    const int n_cols = 2 + 2 * R;
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)n_rep * n_cols) return;
    const int b = (int)(gid / n_cols), col = (int)(gid - (int64_t)b * n_cols);
    const int g = b / reps_per_group, bl = b - g * reps_per_group;
    const int cg = col / kBsOutPitch, lc = col - cg * kBsOutPitch;
    const size_t per = (size_t)kBsMaxRepBlocks * 8 * kBsOutPitch;
    const double* p = partial + ((size_t)cg * n_rep_groups + g) * n_partials * per + (size_t)bl * kBsOutPitch + lc;
    double s = 0.0;
    int i = 0;
    for (; i + 8 <= n_partials; i += 8) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = p[(size_t)(i + u) * per];
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; i < n_partials; ++i) s += p[(size_t)i * per];
    double* dst = acc + (int64_t)b * acc_rep_stride;
    if (col < 2) {
        dst[col] += s;
    } else if (col < 2 + R) {
        dst[col] += unscaled ? s : s * kLegAlpha[col - 2];
    } else {
        const double al = unscaled ? 1.0 : kLegAlpha[col - 2 - R];
        dst[col] += s * (al * al);
    }
}

// Multiplicities of the rows of block blockIdx.x in replicate blockIdx.y: the draws j in [cum[p], cum[p+1]) of
// resample_indices_kernel (same counter, same key, same mapping), histogrammed as packed bytes in shared memory.
__global__ void __launch_bounds__(1024)
resample_counts_kernel(uint64_t seed, uint64_t stream_id, int64_t n_rows, int n_blocks,
                       const int64_t* __restrict__ block_cum, uint8_t* __restrict__ counts, int64_t counts_stride,
                       uint32_t rep_offset) {
    extern __shared__ __align__(16) uint32_t hist[];
    const int p = blockIdx.x, rep = blockIdx.y;
    const int64_t* cum = block_cum + (int64_t)rep * (n_blocks + 1);
    const int64_t j_lo = __ldg(cum + p), j_hi = __ldg(cum + p + 1);
    const int64_t lo = (int64_t)p * n_rows / n_blocks;
    const int64_t size = (int64_t)(p + 1) * n_rows / n_blocks - lo;
    const int words = (int)((size + 3) >> 2);
    for (int i = threadIdx.x; i < words; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    for (int64_t g = (j_lo >> 1) + threadIdx.x; 2 * g < j_hi; g += blockDim.x) {
        uint32_t c[4] = {(uint32_t)g, (uint32_t)(g >> 32), rep_offset + (uint32_t)rep, (uint32_t)stream_id};
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32) ^ (uint32_t)(stream_id >> 32));
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int64_t j = 2 * g + h;
            if (j >= j_lo && j < j_hi) {
                const uint64_t r = ((uint64_t)c[2 * h + 1] << 32) | c[2 * h];
                const uint32_t local = (uint32_t)__umul64hi(r, (uint64_t)size);
                atomicAdd(&hist[local >> 2], 1u << (8 * (local & 3u)));
            }
        }
    }
    __syncthreads();
    uint8_t* out = counts + (int64_t)rep * counts_stride + lo;
    const uint8_t* src = reinterpret_cast<const uint8_t*>(hist);
    // 16-byte stores where the destination is aligned, bytes at the two ends (neighbouring blocks own those words)
    const int64_t to_aligned = (int64_t)((16 - (reinterpret_cast<uintptr_t>(out) & 15)) & 15);
    const int head = (int)(size < to_aligned ? size : to_aligned);
    for (int i = threadIdx.x; i < head; i += blockDim.x) out[i] = src[i];
    const int n_vec = (int)((size - head) >> 4);
    for (int i = threadIdx.x; i < n_vec; i += blockDim.x) {
        uint4 v;
        const uint8_t* s8 = src + head + 16 * i;
        if ((head & 3) == 0) {
            const uint32_t* s32 = reinterpret_cast<const uint32_t*>(s8);
            v = make_uint4(s32[0], s32[1], s32[2], s32[3]);
        } else {
            uint32_t t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                t[u] = (uint32_t)s8[4 * u] | ((uint32_t)s8[4 * u + 1] << 8) | ((uint32_t)s8[4 * u + 2] << 16) |
                       ((uint32_t)s8[4 * u + 3] << 24);
            v = make_uint4(t[0], t[1], t[2], t[3]);
        }
        *reinterpret_cast<uint4*>(out + head + 16 * i) = v;
    }
    for (int64_t i = head + 16 * (int64_t)n_vec + threadIdx.x; i < size; i += blockDim.x) out[i] = src[i];
}

// replicates per group: as few groups as possible, evenly filled, whole blocks of 8
void weighted_groups(int n_rep, int* n_groups, int* reps_per_group) {
    const int nbb = (n_rep + 7) / 8;
    const int g = (nbb + kBsMaxRepBlocks - 1) / kBsMaxRepBlocks;
    *n_groups = g;
    *reps_per_group = 8 * ((nbb + g - 1) / g);
}

int weighted_grid_x(int64_t n_rows) {
    const int64_t tiles = (n_rows + kBsTile - 1) / kBsTile;
    const int64_t sms = sm_count();
    return (int)(tiles < sms ? (tiles > 0 ? tiles : 1) : sms);
}

}  // namespace

}  // namespace mlmcb200

using namespace mlmcb200;

extern "C" int32_t mlmcb200_resample_counts_block_rows(void) { return kCountCapRows; }

extern "C" int mlmcb200_resample_counts(uint64_t seed, uint64_t stream_id, int64_t n_rows, int64_t max_draws,
                                        int32_t n_rep, int32_t rep_offset, int32_t n_blocks, const int64_t* block_cum,
                                        uint8_t* counts, int64_t counts_stride, void* stream) {
    MB_REQUIRE(n_rows >= 1 && n_rows <= 0x7fffffffLL && n_rep >= 1 && n_rep <= 65535 && rep_offset >= 0 &&
                   counts != nullptr && block_cum != nullptr && n_blocks >= 1 && n_blocks <= n_rows,
               "resample_counts: bad arguments (n_rows=%lld n_rep=%d n_blocks=%d)", (long long)n_rows, n_rep, n_blocks);
    MB_REQUIRE((n_rows + n_blocks - 1) / n_blocks <= kCountCapRows,
               "resample_counts: row blocks of more than %d rows (n_rows=%lld, n_blocks=%d)", kCountCapRows,
               (long long)n_rows, n_blocks);
    MB_REQUIRE(counts_stride >= n_rows, "resample_counts: counts_stride=%lld < n_rows", (long long)counts_stride);
    // byte counters: a row drawn 256 times would carry into its neighbour; with at most 8 draws per row on average the
    // chance of that is below 1e-250 per row
    MB_REQUIRE(max_draws >= 0 && max_draws <= 8 * n_rows, "resample_counts: more than 8 draws per row (%lld of %lld)",
               (long long)max_draws, (long long)n_rows);
    const int64_t block_rows = (n_rows + n_blocks - 1) / n_blocks;
    const size_t smem = (size_t)((block_rows + 3) / 4) * 4 + 16;
    static thread_local int attr_dev = -1;
    int dev = 0;
    MB_CUDA_OK(cudaGetDevice(&dev));
    if (attr_dev != dev) {
        MB_CUDA_OK(cudaFuncSetAttribute(resample_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        kCountCapRows + 16));
        attr_dev = dev;
    }
    resample_counts_kernel<<<dim3((unsigned)n_blocks, (unsigned)n_rep), 1024, smem, (cudaStream_t)stream>>>(
        seed, stream_id, n_rows, n_blocks, block_cum, counts, counts_stride, (uint32_t)rep_offset);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}

// column groups of 104: [ok, rm, d_0 .. d_{R-1}, d_0^2 .. d_{R-1}^2]
static int weighted_col_groups(int size) { return (2 + 2 * size + kBsOutPitch - 1) / kBsOutPitch; }

extern "C" int64_t mlmcb200_moments_weighted_workspace_bytes(int64_t n_rows, int32_t n_rep, int32_t size) {
    if (n_rows < 0 || n_rep < 1 || size < 1) return -1;
    int groups = 0, per = 0;
    weighted_groups(n_rep, &groups, &per);
    return (int64_t)weighted_col_groups(size) * groups * weighted_grid_x(n_rows) * (kBsMaxRepBlocks * 8 * kBsOutPitch) *
           (int64_t)sizeof(double);
}

// the limit of the gather kernel for scalar quantities (one pass of 104 columns holds 51 moments, wider bases take
// ceil((2 + 2 R) / 104) passes)
extern "C" int32_t mlmcb200_moments_weighted_max_size(void) { return 226; }

extern "C" int mlmcb200_moments_accumulate_weighted(const mlmcb200_basis_t* basis, const double* pairs, int64_t n_rows,
                                                    int64_t stride_n, int32_t has_coarse, const uint8_t* counts,
                                                    int64_t counts_stride, int32_t n_rep, double* acc,
                                                    int64_t acc_rep_stride, void* workspace, int64_t workspace_bytes,
                                                    void* stream) {
    if (check_basis(basis) != 0) return -1;
    MB_REQUIRE(basis->kind != MLMCB200_RAW && basis->size <= mlmcb200_moments_weighted_max_size(),
               "moments_accumulate_weighted: Legendre / Monomial / Fourier bases of at most %d moments (kind=%d size=%d)",
               mlmcb200_moments_weighted_max_size(), basis->kind, basis->size);
    MB_REQUIRE(n_rows >= 0 && n_rep >= 1 && acc != nullptr && workspace != nullptr && counts != nullptr,
               "moments_accumulate_weighted: bad arguments");
    MB_REQUIRE(acc_rep_stride >= 2 + 2 * (int64_t)basis->size, "moments_accumulate_weighted: replicate stride too small");
    MB_REQUIRE(counts_stride >= n_rows && counts_stride % 16 == 0 && (reinterpret_cast<uintptr_t>(counts) & 15) == 0,
               "moments_accumulate_weighted: counts rows must be 16-byte aligned and at least n_rows long");
    MB_REQUIRE(has_coarse ? (stride_n == 2 && (reinterpret_cast<uintptr_t>(pairs) & 15) == 0) : stride_n >= 1,
               "moments_accumulate_weighted: scalar quantity in storage order expected ((fine, coarse) pairs 16-byte "
               "aligned, stride_n=%lld)", (long long)stride_n);
    if (n_rows == 0) return 0;
    MB_REQUIRE(pairs != nullptr, "moments_accumulate_weighted: null pairs");
    MB_REQUIRE(workspace_bytes >= mlmcb200_moments_weighted_workspace_bytes(n_rows, n_rep, basis->size),
               "moments_accumulate_weighted: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    WeightedArgs a;
    a.basis = *basis;
    a.pairs = pairs;
    a.n = n_rows;
    a.stride_n = stride_n;
    a.counts = counts;
    a.counts_stride = counts_stride;
    a.n_rep = n_rep;
    int groups = 0;
    weighted_groups(n_rep, &groups, &a.reps_per_group);
    a.partial = static_cast<double*>(workspace);
    const size_t smem = (size_t)(kBsTile * kBsLD + 8) * sizeof(double) + (size_t)kBsMaxRepBlocks * 8 * kBsPitch;
    const int col_groups = weighted_col_groups(basis->size);
    const dim3 grid((unsigned)weighted_grid_x(n_rows), (unsigned)groups, (unsigned)col_groups);
#define MB_LAUNCH(C, M)                                                                                          \
    {                                                                                                            \
        MB_CUDA_OK(cudaFuncSetAttribute(weighted_moments_kernel<C, M>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                        (int)smem));                                                             \
        weighted_moments_kernel<C, M><<<grid, kBsThreads, smem, st>>>(a);                                        \
    }
    if (has_coarse) {
        if (col_groups > 1) MB_LAUNCH(true, true) else MB_LAUNCH(true, false)
    } else {
        if (col_groups > 1) MB_LAUNCH(false, true) else MB_LAUNCH(false, false)
    }
#undef MB_LAUNCH
    MB_CUDA_OK(cudaGetLastError());
    const int64_t outs = (int64_t)n_rep * (2 + 2 * basis->size);
    weighted_finish_kernel<<<(unsigned)((outs + 127) / 128), 128, 0, st>>>(a.partial, (int)grid.x, n_rep,
                                                                           a.reps_per_group, groups, basis->size,
                                                                           basis->kind != MLMCB200_LEGENDRE ? 1 : 0, acc,
                                                                           acc_rep_stride);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}
