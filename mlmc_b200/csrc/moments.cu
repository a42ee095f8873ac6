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
    double* partial;       // [gridDim.y][2 + 2K]
    int64_t partial_stride;
};

// ---- per-moment accumulate: a1 += sum_s d_s, a2 += sum_s d_s^2 over the S samples held by this thread.
// Pairwise tree for the sum and two interleaved FMA chains for the squares keep the dependency depth at
// log2(S) / S/2 instead of S (the FP64 pipe issues in order; a serial chain would stall it).
template <bool COARSE, int S>
__device__ __forceinline__ void accumulate(double* __restrict__ col_sum, double* __restrict__ col_sq,
                                           const double (&vf)[S], const double (&vc)[S]) {
    static_assert(S % 2 == 0, "S must be even");
    double d[S];
#pragma unroll
    for (int s = 0; s < S; ++s) d[s] = COARSE ? vf[s] - vc[s] : vf[s];
    double qa = *col_sq, qb = d[1] * d[1];
    qa = fma(d[0], d[0], qa);
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
    *col_sum += d[0];
    *col_sq = qa + qb;
}

// raw (fine, coarse) values of the S samples of one tile owned by this thread; out-of-range samples read NaN
template <bool COARSE, int S>
__device__ __forceinline__ void load_tile(const MomentsArgs& a, const double* base_f, int64_t n0, int TN, bool active,
                                          double (&xf)[S], double (&xc)[S]) {
    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
#pragma unroll
    for (int s = 0; s < S; ++s) {
        const int64_t n = n0 + (int64_t)s * TN;
        xf[s] = qnan;
        xc[s] = qnan;
        if (active && n < a.n) {
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

template <int KIND, bool COARSE, bool LOG, int S>
__global__ void __launch_bounds__(kThreads)
moments_acc_kernel(const MomentsArgs a) {
    extern __shared__ double sm[];
    const int T = kThreads;
    const int tid = threadIdx.x;
    const int R = a.basis.size;
    const int M = a.n_comp;
    double* const sum_col = sm + tid;                       // element r at sum_col[r * T]
    double* const sq_col = sm + (size_t)R * T + tid;

    for (int r = 0; r < R; ++r) {
        sum_col[r * T] = 0.0;
        sq_col[r * T] = 0.0;
    }

    // thread -> (component, sample lane)
    int m, tn, TN;
    bool active;
    if (M >= T) {
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
    const bool count_here = (blockIdx.x == 0) && (M >= T ? tid == 0 : m == 0);

    unsigned cnt_ok = 0, cnt_rm = 0;

    // software pipeline: the raw values of the NEXT tile are in flight while the current tile is reduced
    double xf[S], xc[S];
    load_tile<COARSE, S>(a, base_f, (int64_t)blockIdx.y * tile_n + tn, TN, active, xf, xc);

    for (int64_t tile = blockIdx.y; tile < n_tiles; tile += gridDim.y) {
        double tf[S], tc[S];
        bool ok[S];
        const int64_t n0 = tile * tile_n + tn;
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
            bool good;
            if (a.valid != nullptr) {
                good = in && a.valid[n] != 0;
            } else {
                good = in && moments_finite(a.basis, tf[s]) && (!COARSE || moments_finite(a.basis, tc[s]));
            }
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
            load_tile<COARSE, S>(a, base_f, (tile + gridDim.y) * tile_n + tn, TN, active, xf, xc);

        if (KIND == MLMCB200_RAW) {
            accumulate<COARSE, S>(sum_col, sq_col, tf, tc);
        } else if (KIND == MLMCB200_LEGENDRE || KIND == MLMCB200_MONOMIAL) {
            // Two ping-pong register sets hold moment k of the S samples: even k in (fb, cb), odd k in (fa, ca).
            //   Legendre (monic): W_k = t W_{k-1} - e_k W_{k-2}, the e_k product in place on the older set
            //   Monomial        : t^k = t^{k-1} t
            // Software pipeline: the reduction of moment k is issued AFTER the recurrence of moment k+1, so the
            // dependent tail of one (add tree, smem update) overlaps the independent FP64 work of the other.
            double fa[S], fb[S], ca[S], cb[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                fb[s] = cb[s] = ok[s] ? 1.0 : 0.0;
                fa[s] = tf[s];
                ca[s] = tc[s];
            }
            accumulate<COARSE, S>(sum_col, sq_col, fb, cb);                       // moment 0
#define MB_REC(DST_F, DST_C, SRC_F, SRC_C, E)                                   \
    _Pragma("unroll") for (int s = 0; s < S; ++s) {                              \
        if (KIND == MLMCB200_LEGENDRE) {                                         \
            DST_F[s] *= (E);                                                     \
            if (COARSE) DST_C[s] *= (E);                                         \
        }                                                                        \
    }                                                                            \
    _Pragma("unroll") for (int s = 0; s < S; ++s) {                              \
        if (KIND == MLMCB200_LEGENDRE) {                                         \
            DST_F[s] = fma(tf[s], SRC_F[s], -DST_F[s]);                          \
            if (COARSE) DST_C[s] = fma(tc[s], SRC_C[s], -DST_C[s]);              \
        } else {                                                                 \
            DST_F[s] = SRC_F[s] * tf[s];                                         \
            if (COARSE) DST_C[s] = SRC_C[s] * tc[s];                             \
        }                                                                        \
    }
#define MB_ACC(K, VF, VC) accumulate<COARSE, S>(sum_col + (K) * T, sq_col + (K) * T, VF, VC)
            int i = 2;                                                            // next moment to generate
            if (R > 5) {
                double e0 = kLegCoef[2], e1 = kLegCoef[3], e2 = kLegCoef[4], e3 = kLegCoef[5];
                for (; i + 3 < R; i += 4) {
                    // coefficients of the NEXT group are fetched now (the table is padded past MAX_MOMENTS)
                    const double n0 = kLegCoef[i + 4], n1 = kLegCoef[i + 5], n2 = kLegCoef[i + 6], n3 = kLegCoef[i + 7];
                    MB_REC(fb, cb, fa, ca, e0)
                    MB_ACC(i - 1, fa, ca);
                    MB_REC(fa, ca, fb, cb, e1)
                    MB_ACC(i, fb, cb);
                    MB_REC(fb, cb, fa, ca, e2)
                    MB_ACC(i + 1, fa, ca);
                    MB_REC(fa, ca, fb, cb, e3)
                    MB_ACC(i + 2, fb, cb);
                    e0 = n0; e1 = n1; e2 = n2; e3 = n3;
                }
            }
            for (; i < R; ++i) {                                                 // remainder, one moment at a time
                const double e = kLegCoef[i];
                if ((i & 1) == 0) {
                    MB_REC(fb, cb, fa, ca, e)
                    MB_ACC(i - 1, fa, ca);
                } else {
                    MB_REC(fa, ca, fb, cb, e)
                    MB_ACC(i - 1, fb, cb);
                }
            }
            if (R > 1) {                                                          // the last moment is still pending
                if ((R - 1) & 1) MB_ACC(R - 1, fa, ca);
                else MB_ACC(R - 1, fb, cb);
            }
#undef MB_REC
#undef MB_ACC
        } else {  // FOURIER: columns 1, cos t, sin t, cos 2t, sin 2t, ... by exact-angle rotation
            double cf1[S], sf1[S], cc1[S], sc1[S], cfk[S], sfk[S], cck[S], sck[S];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                cfk[s] = cck[s] = ok[s] ? 1.0 : 0.0;
                cf1[s] = sf1[s] = cc1[s] = sc1[s] = 0.0;
                if (ok[s] && R > 1) {
                    sincos(tf[s], &sf1[s], &cf1[s]);
                    if (COARSE) sincos(tc[s], &sc1[s], &cc1[s]);
                }
            }
            accumulate<COARSE, S>(sum_col, sq_col, cfk, cck);
#pragma unroll
            for (int s = 0; s < S; ++s) {
                cfk[s] = cf1[s];
                sfk[s] = sf1[s];
                cck[s] = cc1[s];
                sck[s] = sc1[s];
            }
            for (int i = 1; i < R; i += 2) {
                accumulate<COARSE, S>(sum_col + i * T, sq_col + i * T, cfk, cck);
                if (i + 1 < R) accumulate<COARSE, S>(sum_col + (i + 1) * T, sq_col + (i + 1) * T, sfk, sck);
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

    // ---------------- block epilogue ----------------
    __syncthreads();
    double* const out = a.partial + (int64_t)blockIdx.y * a.partial_stride;
    const int64_t K = (int64_t)M * R;
    if (M >= T) {
        if (active) {
            for (int r = 0; r < R; ++r) {
                const double al = (KIND == MLMCB200_LEGENDRE) ? kLegAlpha[r] : 1.0;
                out[2 + (int64_t)m * R + r] = sum_col[r * T] * al;
                out[2 + K + (int64_t)m * R + r] = sq_col[r * T] * (al * al);
            }
        }
    } else {
        const int warp = tid >> 5, lane = tid & 31, n_warps = T >> 5;
        for (int k = warp; k < (int)K; k += n_warps) {
            const int mm = k / R, r = k - mm * R;
            double s1 = 0.0, s2 = 0.0;
            for (int j = lane; j < TN; j += 32) {
                s1 += sm[r * T + j * M + mm];
                s2 += sm[(size_t)R * T + r * T + j * M + mm];
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
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int n_partials, int64_t stride,
                                       int64_t len, double* __restrict__ acc, int warp_per_output) {
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (warp_per_output) {
        const int64_t j = gid >> 5;
        const int lane = threadIdx.x & 31;
        if (j >= len) return;
        double s = 0.0;
        for (int b = lane; b < n_partials; b += 32) s += partial[(int64_t)b * stride + j];
        s = warp_sum(s);
        if (lane == 0) acc[j] += s;
    } else {
        if (gid >= len) return;
        double s = 0.0;
        for (int b = 0; b < n_partials; ++b) s += partial[(int64_t)b * stride + gid];
        acc[gid] += s;
    }
}

__global__ void sample_mask_kernel(const mlmcb200_basis_t basis, const double* __restrict__ pairs, int64_t n,
                                   int n_comp, int64_t stride_n, int64_t stride_side, int64_t stride_m,
                                   int n_sides, uint8_t* __restrict__ valid) {
    // one warp per sample; lanes stride over the components of both sides
    const int64_t sample = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (sample >= n) return;
    bool good = true;
    const double* row = pairs + sample * stride_n;
    for (int side = 0; side < n_sides; ++side) {
        for (int mm = lane; mm < n_comp; mm += 32) {
            double x = __ldg(row + side * stride_side + (int64_t)mm * stride_m);
            double t = basis.kind == MLMCB200_RAW ? x : map_to_ref(basis, x);
            good = good && moments_finite(basis, t);
        }
    }
    good = __all_sync(0xffffffffu, good);
    if (lane == 0) valid[sample] = good ? 1 : 0;
}

__global__ void finalize_levels_kernel(const double* __restrict__ acc, int64_t acc_stride, int n_levels, int64_t K,
                                       double* l_means, double* l_vars, double* mean, double* var) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
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

}  // namespace

int launch_reduce_partials(const double* partial, int n_partials, int64_t stride, int64_t len, double* acc,
                           cudaStream_t st) {
    const int threads = 256;
    const int wpo = n_partials >= 32 && len <= (1 << 20) ? 1 : 0;
    const int64_t total = wpo ? len * 32 : len;
    reduce_partials_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, st>>>(partial, n_partials,
                                                                                            stride, len, acc, wpo);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}

namespace {

struct Plan {
    dim3 grid;
    size_t smem;
    int samples_per_thread;
};

int plan_moments(int kind, int size, int n_comp, int64_t n, Plan* p) {
    const int S = (kind == MLMCB200_FOURIER) ? 4 : 8;
    const size_t smem = (size_t)2 * size * kThreads * sizeof(double);
    if (smem > 227u * 1024u) {
        set_error("moments: size %d needs %zu B of shared memory per CTA (max %u)", size, smem, 227u * 1024u);
        return -1;
    }
    int ctas_per_sm = (int)((227u * 1024u) / (smem + 1024));
    if (ctas_per_sm > 4) ctas_per_sm = 4;
    if (ctas_per_sm < 1) ctas_per_sm = 1;
    const int gx = n_comp >= kThreads ? (n_comp + kThreads - 1) / kThreads : 1;
    int gy = (sm_count() * ctas_per_sm + gx - 1) / gx;
    if (n >= 0) {
        const int TN = n_comp >= kThreads ? 1 : kThreads / n_comp;
        const int64_t tiles = (n + (int64_t)S * TN - 1) / ((int64_t)S * TN);
        if (tiles < gy) gy = (int)(tiles > 0 ? tiles : 1);
    }
    p->grid = dim3(gx, gy, 1);
    p->smem = smem;
    p->samples_per_thread = S;
    return 0;
}

template <int KIND, bool COARSE, bool LOG, int S>
int launch_moments(const MomentsArgs& a, Plan p, cudaStream_t st) {
    auto kern = moments_acc_kernel<KIND, COARSE, LOG, S>;
    // resident CTAs per SM for this variant and shared-memory size (registers may bind before shared memory);
    // the sample-partition dimension of the grid is sized to exactly one resident wave
    static thread_local size_t cached_smem = 0;
    static thread_local int cached_ctas = 0;
    if (cached_smem != p.smem || cached_ctas == 0) {
        MB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
        int ctas = 0;
        MB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas, kern, kThreads, p.smem));
        cached_ctas = ctas > 0 ? ctas : 1;
        cached_smem = p.smem;
    }
    const unsigned wave = (unsigned)((sm_count() * cached_ctas + p.grid.x - 1) / p.grid.x);
    if (p.grid.y > wave) p.grid.y = wave;
    kern<<<p.grid, kThreads, p.smem, st>>>(a);
    MB_CUDA_OK(cudaGetLastError());
    return (int)p.grid.y;
}

}  // namespace

}  // namespace mlmcb200

using namespace mlmcb200;

extern "C" int64_t mlmcb200_moments_workspace_bytes(int32_t size, int32_t n_comp) {
    Plan p;
    if (size < 1 || n_comp < 1) return -1;
    if (plan_moments(MLMCB200_LEGENDRE, size, n_comp, -1, &p) != 0) return -1;
    return (int64_t)p.grid.y * (2 + 2 * (int64_t)size * n_comp) * (int64_t)sizeof(double);
}

extern "C" int mlmcb200_moments_accumulate(const mlmcb200_basis_t* basis, const double* pairs, int64_t n,
                                           int32_t n_comp, int64_t stride_n, int64_t stride_side,
                                           int64_t stride_m, int32_t has_coarse, const uint8_t* valid,
                                           double* acc, void* workspace, int64_t workspace_bytes, void* stream) {
    if (check_basis(basis) != 0) return -1;
    MB_REQUIRE(n >= 0 && n_comp >= 1, "moments_accumulate: bad n=%lld n_comp=%d", (long long)n, n_comp);
    MB_REQUIRE(acc != nullptr && workspace != nullptr, "moments_accumulate: null acc/workspace");
    MB_REQUIRE(n_comp == 1 || valid != nullptr, "moments_accumulate: n_comp > 1 needs the sample mask");
    if (n == 0) return 0;
    MB_REQUIRE(pairs != nullptr, "moments_accumulate: null pairs");
    cudaStream_t st = (cudaStream_t)stream;
    Plan p;
    if (plan_moments(basis->kind, basis->size, n_comp, n, &p) != 0) return -1;
    const int64_t K = (int64_t)basis->size * n_comp;
    const int64_t stride = 2 + 2 * K;
    MB_REQUIRE(workspace_bytes >= (int64_t)p.grid.y * stride * 8, "moments_accumulate: workspace too small");

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

    int rc = -1;     // > 0: number of partial vectors written
#define MB_DISPATCH(KIND, S)                                                                               \
    rc = basis->is_log                                                                                      \
             ? (has_coarse ? launch_moments<KIND, true, true, S>(a, p, st)                                  \
                           : launch_moments<KIND, false, true, S>(a, p, st))                                \
             : (has_coarse ? launch_moments<KIND, true, false, S>(a, p, st)                                 \
                           : launch_moments<KIND, false, false, S>(a, p, st))
    switch (basis->kind) {
        case MLMCB200_RAW:
            rc = has_coarse ? launch_moments<MLMCB200_RAW, true, false, 8>(a, p, st)
                            : launch_moments<MLMCB200_RAW, false, false, 8>(a, p, st);
            break;
        case MLMCB200_LEGENDRE: MB_DISPATCH(MLMCB200_LEGENDRE, 8); break;
        case MLMCB200_MONOMIAL: MB_DISPATCH(MLMCB200_MONOMIAL, 8); break;
        case MLMCB200_FOURIER: MB_DISPATCH(MLMCB200_FOURIER, 4); break;
    }
#undef MB_DISPATCH
    if (rc <= 0) return rc < 0 ? rc : -1;
    return launch_reduce_partials(a.partial, rc, stride, stride, acc, st);
}

extern "C" int mlmcb200_sample_mask(const mlmcb200_basis_t* basis, const double* pairs, int64_t n, int32_t n_comp,
                                    int64_t stride_n, int64_t stride_side, int64_t stride_m, int32_t has_coarse,
                                    uint8_t* valid, void* stream) {
    if (check_basis(basis) != 0) return -1;
    MB_REQUIRE(n >= 0 && n_comp >= 1 && valid != nullptr, "sample_mask: bad arguments");
    if (n == 0) return 0;
    const int threads = 256;
    const int64_t blocks = (n * 32 + threads - 1) / threads;
    sample_mask_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(
        *basis, pairs, n, n_comp, stride_n, stride_side, stride_m, has_coarse ? 2 : 1, valid);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int mlmcb200_finalize_levels(const double* acc, int64_t acc_stride, int32_t n_levels, int64_t K,
                                        double* l_means, double* l_vars, double* mean, double* var, void* stream) {
    MB_REQUIRE(acc != nullptr && n_levels >= 1 && K >= 1 && acc_stride >= 2 + 2 * K, "finalize_levels: bad arguments");
    const int threads = 128;
    finalize_levels_kernel<<<(unsigned)((K + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        acc, acc_stride, n_levels, K, l_means, l_vars, mean, var);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}
