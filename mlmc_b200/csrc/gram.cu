// Moment-covariance level sums on FP64 tensor-core tiles (DMMA m8n8k4), sm_100a.
//
// Replaces estimate_mean over a `covariance` quantity (mlmc/quantity/quantity_estimate.py:131-147 + :43-65).
// The reference materialises per-sample outer products [R, R, n, 2]; here a CTA turns a tile of samples into the
// basis tables Phi_f, Phi_c in SHARED memory only (never in HBM) and contracts them with mma.sync f64:
//     sum_n d_ij    = Phi_f^T Phi_f - Phi_c^T Phi_c                                   (d_ij = f_i f_j - c_i c_j)
//     sum_n d_ij^2  = (D.D)^T (F.F) + 2 (D.C)^T (F.D) + (C.C)^T (D.D),   D = Phi_f - Phi_c
// (second line: d_ij = D_i f_j + c_i D_j, expanded; "." = elementwise).  Every product is symmetric in (i, j),
// so only 8x8 blocks on or above the diagonal are computed and mirrored on output.
//
// tcgen05.mma has no f64 kind; the FP64 tensor path of sm_100 is the warp-level DMMA.8x8x4
// (mma.sync.aligned.m8n8k4.row.col.f64), SURVEY.md section 7 "hard parts".
//
// Fragment layout of m8n8k4 (lane l): A[row l/4][k l%4], B[k l%4][col l/4], C[row l/4][cols 2(l%4), 2(l%4)+1].
// With A = Phi^T and B = Phi (k = sample), BOTH operand fragments of moment block b are the same smem element
// Phi[n0 + l%4][8b + l/4]; a leading dimension LD = 4 (mod 8) doubles makes that access conflict-free.
#include <stdlib.h>
#include "common.cuh"

namespace mlmcb200 {
namespace {

constexpr int kWarps = 16;
constexpr int kThreadsGram = kWarps * 32;
constexpr int kMaxSlots = 4;     // tasks per warp (array bound; the plans use 2 ... 4)
// Leading dimension of the covariance tiles: a compile-time constant (template parameter; 4 mod 8, >= 8 nb + 3 for the
// row skew) so that every fragment address of a tile is "slot base register + immediate".  Three sizes: up to 4 / 7 / 13
// blocks of 8 moments -- a narrower tile holds more samples (128 -> 224 -> 384 with two arrays), which spreads the
// produce phase over more warps and the per-tile costs (two barriers, task dispatch) over more DMMAs.
constexpr int kLDSmall = 36, kLDMid = 60, kLDWide = 108;
constexpr int kGramMaxMoments = 104;
inline int gram_ld(int nb) { return nb <= 4 ? kLDSmall : nb <= 7 ? kLDMid : kLDWide; }

struct GramPlan {
    int nb;          // 8x8 blocks per side
    int gs;          // blocks per group (1 or 2)
    int ld;          // smem leading dimension (doubles)
    int ns;          // samples per tile (multiple of 4)
    int n_tasks[kWarps];
    unsigned char gi[kWarps][kMaxSlots];
    unsigned char gj[kWarps][kMaxSlots];
    unsigned char mk[kWarps][kMaxSlots];   // which of the task's gs x gs blocks this slot owns (bit u*gs+v)
};

struct GramArgs {
    mlmcb200_basis_t basis;
    const double* pairs;     // component blockIdx.y starts at pairs + blockIdx.y * stride_m
    int64_t n, stride_n, stride_side, stride_m;
    const uint8_t* valid;    // vector quantities: sample mask shared by the components (mlmcb200_sample_mask), or null
    double* partial;         // [gridDim.y][gridDim.x][2 + 2 R R]
    int64_t partial_stride;
    GramPlan plan;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// One basis row (R values, zero-padded to 8*nb) into shared memory.  Zero row for a dropped sample.
__device__ __forceinline__ void write_row(const mlmcb200_basis_t& b, double t, bool good, double* row, int r_pad) {
    const int R = b.size;
    if (!good) {
        for (int i = 0; i < r_pad; ++i) row[i] = 0.0;
        return;
    }
    if (b.kind == MLMCB200_RAW) {
        row[0] = t;
    } else if (b.kind == MLMCB200_FOURIER) {
        row[0] = 1.0;
        if (R > 1) {
            double s1, c1;
            sincos(t, &s1, &c1);
            double ck = c1, sk = s1;
            for (int i = 1; i < R; i += 2) {
                row[i] = ck;
                if (i + 1 < R) row[i + 1] = sk;
                const double nc = fma(ck, c1, -(sk * s1));
                sk = fma(sk, c1, ck * s1);
                ck = nc;
            }
        }
    } else {
        double p0 = 1.0, p1 = t;
        row[0] = p0;
        if (R > 1) row[1] = p1;
        if (b.kind == MLMCB200_LEGENDRE) {
            // MONIC recurrence W_i = t W_{i-1} - e_i W_{i-2} (P_i = g_i W_i, applied to the sums in the epilogue): two FP64
            // instructions per element instead of three.  Four steps per iteration, coefficient loads hoisted.
            // W_i = fma(t, W_{i-1}, -(e_i W_{i-2})): the product is off the dependent chain (one DFMA latency per step).
            int i = 2;
            for (; i + 3 < R; i += 4) {
                const double e0 = kLegCoef[i], e1 = kLegCoef[i + 1], e2 = kLegCoef[i + 2], e3 = kLegCoef[i + 3];
                const double q0 = fma(t, p1, -(e0 * p0));
                const double q1 = fma(t, q0, -(e1 * p1));
                const double q2 = fma(t, q1, -(e2 * q0));
                const double q3 = fma(t, q2, -(e3 * q1));
                row[i] = q0;
                row[i + 1] = q1;
                row[i + 2] = q2;
                row[i + 3] = q3;
                p0 = q2;
                p1 = q3;
            }
            for (; i < R; ++i) {
                const double p2 = fma(t, p1, -(kLegCoef[i] * p0));
                row[i] = p2;
                p0 = p1;
                p1 = p2;
            }
        } else {
            for (int i = 2; i < R; ++i) {
                const double p2 = p1 * t;
                row[i] = p2;
                p0 = p1;
                p1 = p2;
            }
        }
    }
    for (int i = R; i < r_pad; ++i) row[i] = 0.0;
}

// Fine + coarse levels: the tile holds S = phi(f) + phi(c) and D = phi(f) - phi(c) instead of phi(f), phi(c)
// (d_ij = f_i f_j - c_i c_j = (D_i S_j + S_i D_j) / 2: no cancellation between two Gram matrices, see slot_mma).
// One thread runs both recurrences of a sample (two independent chains; a lane-pair variant that exchanged every
// element by shuffle was 4 % slower).  Zero rows for a dropped sample.
// ONLY_D (Gram of the differences): the tile is the single array D.
template <bool ONLY_D>
__device__ __forceinline__ void write_rows_sd(const mlmcb200_basis_t& b, double tf, double tc, bool good, double* row_s,
                                              double* row_d, int r_pad) {
    const int R = b.size;
    if (!good) {
        for (int i = 0; i < r_pad; ++i) {
            if (!ONLY_D) row_s[i] = 0.0;
            row_d[i] = 0.0;
        }
        return;
    }
#define MB_PUT(I, F, C)                         \
    {                                           \
        if (!ONLY_D) row_s[I] = (F) + (C);      \
        row_d[I] = (F) - (C);                   \
    }
    if (b.kind == MLMCB200_RAW) {
        MB_PUT(0, tf, tc)
    } else if (b.kind == MLMCB200_FOURIER) {
        MB_PUT(0, 1.0, 1.0)
        if (R > 1) {
            double sf1, cf1, sc1, cc1;
            sincos(tf, &sf1, &cf1);
            sincos(tc, &sc1, &cc1);
            double cfk = cf1, sfk = sf1, cck = cc1, sck = sc1;
            for (int i = 1; i < R; i += 2) {
                MB_PUT(i, cfk, cck)
                if (i + 1 < R) MB_PUT(i + 1, sfk, sck)
                const double nf = fma(cfk, cf1, -(sfk * sf1));
                sfk = fma(sfk, cf1, cfk * sf1);
                cfk = nf;
                const double nc = fma(cck, cc1, -(sck * sc1));
                sck = fma(sck, cc1, cck * sc1);
                cck = nc;
            }
        }
    } else {
        double f0 = 1.0, f1 = tf, c0 = 1.0, c1 = tc;
        MB_PUT(0, f0, c0)
        if (R > 1) MB_PUT(1, f1, c1)
        if (b.kind == MLMCB200_LEGENDRE) {
            // monic recurrence W_i = t W_{i-1} - e_i W_{i-2} for both values, two steps per iteration
            int i = 2;
            for (; i + 1 < R; i += 2) {
                const double e0 = kLegCoef[i], e1 = kLegCoef[i + 1];
                const double qf0 = fma(tf, f1, -(e0 * f0)), qc0 = fma(tc, c1, -(e0 * c0));
                const double qf1 = fma(tf, qf0, -(e1 * f1)), qc1 = fma(tc, qc0, -(e1 * c1));
                MB_PUT(i, qf0, qc0)
                MB_PUT(i + 1, qf1, qc1)
                f0 = qf0;
                f1 = qf1;
                c0 = qc0;
                c1 = qc1;
            }
            if (i < R) {
                const double qf = fma(tf, f1, -(kLegCoef[i] * f0)), qc = fma(tc, c1, -(kLegCoef[i] * c0));
                MB_PUT(i, qf, qc)
            }
        } else {
            for (int i = 2; i < R; ++i) {
                f1 *= tf;
                c1 *= tc;
                MB_PUT(i, f1, c1)
            }
        }
    }
#undef MB_PUT
    for (int i = R; i < r_pad; ++i) {
        if (!ONLY_D) row_s[i] = 0.0;
        row_d[i] = 0.0;
    }
}

// Coefficients of TWO steps of the monic Legendre recurrence at once (gen_tables.py):
//     W_j = (t^2 - A_j) W_{j-2} - B_j W_{j-4}
// The even and the odd moments of a sample are then independent chains, and the producers of a tile split every sample
// over two threads (by parity): twice as many warps advance the tile, which matters because the produce phase is bound
// by the latency of its dependent FP64 chains (one warp per SM sub-partition issued an instruction every ~3.3 cycles),
// not by the FP64 pipe.
static __constant__ double kLegA2[MLMCB200_MAX_MOMENTS] = MLMCB200_LEG_A2_INIT;
static __constant__ double kLegB2[MLMCB200_MAX_MOMENTS] = MLMCB200_LEG_B2_INIT;

// Moments PARITY, PARITY + 2, ... of one sample (Legendre, monic W_j): rows S / D (COARSE), D alone (ONLY_D) or phi(f).
// PARITY is a template parameter, the stores walk their own pointers and the caller reaches this code through
// warp-uniform branches only, so that the moment index j depends on nothing but constants and kernel parameters: the
// compiler keeps it -- and the coefficients it fetches -- in UNIFORM registers (ULDC instead of an indexed LDC per
// coefficient through the MIO queue, which the stores need).  tf = tc = 0 for a dropped sample.
template <bool COARSE, bool ONLY_D, int PARITY>
__device__ __forceinline__ void write_rows_parity(int R, double tf, double tc, bool good, double* row_s, double* row_d,
                                                  int r_pad) {
#define MB_PUT(OFF, F, C)                                   \
    {                                                       \
        if (!COARSE) ps[OFF] = (F);                         \
        else {                                              \
            if (!ONLY_D) ps[OFF] = (F) + (C);               \
            pd[OFF] = (F) - (C);                            \
        }                                                   \
    }
    double* ps = row_s + PARITY;
    double* pd = row_d + PARITY;
    if (PARITY >= R) {                                      // uniform: a basis of one function has no odd moments
        for (int i = PARITY; i < r_pad; i += 2, ps += 2, pd += 2) MB_PUT(0, 0.0, 0.0)
        return;
    }
    // A dropped sample starts from W = 0 (t = 0) and so produces exact zeros: no thread-dependent branch around the
    // loop, which therefore runs converged and may use the uniform datapath.
    const double uf = tf * tf, uc = tc * tc;
    double f0 = 0.0, c0 = 0.0;                              // W_{j-4}
    double f1 = good ? (PARITY ? tf : 1.0) : 0.0, c1 = good ? (PARITY ? tc : 1.0) : 0.0;   // W_{j-2}
    MB_PUT(0, f1, c1)
    ps += 2;
    pd += 2;
    int j = PARITY + 2;
    for (; j + 2 < R; j += 4, ps += 4, pd += 4) {           // two steps per iteration
        const double a0 = kLegA2[j], b0 = kLegB2[j], a1 = kLegA2[j + 2], b1 = kLegB2[j + 2];
        const double qf0 = fma(uf - a0, f1, -(b0 * f0));
        const double qf1 = fma(uf - a1, qf0, -(b1 * f1));
        double qc0 = 0.0, qc1 = 0.0;
        if (COARSE) {
            qc0 = fma(uc - a0, c1, -(b0 * c0));
            qc1 = fma(uc - a1, qc0, -(b1 * c1));
        }
        MB_PUT(0, qf0, qc0)
        MB_PUT(2, qf1, qc1)
        f0 = qf0;
        f1 = qf1;
        c0 = qc0;
        c1 = qc1;
    }
    if (j < R) {
        const double a0 = kLegA2[j], b0 = kLegB2[j];
        const double qf = fma(uf - a0, f1, -(b0 * f0));
        const double qc = COARSE ? fma(uc - a0, c1, -(b0 * c0)) : 0.0;
        MB_PUT(0, qf, qc)
        j += 2;
        ps += 2;
        pd += 2;
    }
    for (; j < r_pad; j += 2, ps += 2, pd += 2) MB_PUT(0, 0.0, 0.0)
#undef MB_PUT
}

// All DMMAs of one task for one 4-sample step.  MASK = the 8x8 blocks of the GS x GS group this slot owns
// (bit u*GS+v).  Products into the same accumulator are issued in separate passes so that consecutive DMMAs are
// independent; only the fragments the mask needs are loaded.
//
// COARSE tiles hold S = phi(f) + phi(c) (ps) and D = phi(f) - phi(c) (pd).  With d_ij = (D_i S_j + S_i D_j) / 2:
//   2 sum d_ij    = (S^T D + D^T S)_ij                               two DMMAs per block, ONE for a diagonal block:
//                                                                    X = S_I^T D_I, the epilogue / reduction forms X + X^T
//   4 sum d_ij^2  = (A^T B + B^T A + 2 E^T E)_ij,  A = D.D, B = S.S, E = D.S   three DMMAs per block, two on the diagonal
//                                                                    (Y = A_I^T B_I + E_I^T E_I, Y + Y^T is the block)
//   MODE 2:         (D^T D)_ij                                       one DMMA per block, no subtraction here
// DIAG: the task lies on the block diagonal (gi == gj): its blocks (u, u) are diagonal blocks.
template <bool COARSE, int MODE, int GS, int MASK, bool DIAG>
__device__ __forceinline__ void slot_mma(double (&am)[GS][GS][2], double (&av)[GS][GS][2], const double* ps,
                                         const double* pd, const int (&off_r)[GS], const int (&off_c)[GS]) {
    double sr[GS], dr[GS], scol[GS], dcol[GS];
#pragma unroll
    for (int u = 0; u < GS; ++u) {
        sr[u] = dr[u] = scol[u] = dcol[u] = 0.0;
        bool row_used = false, col_used = false;
#pragma unroll
        for (int v = 0; v < GS; ++v) {
            row_used = row_used || ((MASK >> (u * GS + v)) & 1);
            col_used = col_used || ((MASK >> (v * GS + u)) & 1);
        }
        if (DIAG) {
            // a task on the block diagonal has the same moment blocks on both sides: one set of fragments (and, with the
            // sums of squares, one set of element-wise products) serves rows and columns
            if (row_used || col_used) {
                if (!(COARSE && MODE == 2)) sr[u] = ps[off_r[u]];
                if (COARSE) dr[u] = pd[off_r[u]];
            }
            scol[u] = sr[u];
            dcol[u] = dr[u];
        } else {
            if (row_used) {
                if (!(COARSE && MODE == 2)) sr[u] = ps[off_r[u]];
                if (COARSE) dr[u] = pd[off_r[u]];
            }
            if (col_used) {
                if (!(COARSE && MODE == 2)) scol[u] = ps[off_c[u]];
                if (COARSE) dcol[u] = pd[off_c[u]];
            }
        }
    }
#define MB_FOR_BLOCKS(COND, BODY)                                               \
    _Pragma("unroll") for (int u = 0; u < GS; ++u) {                            \
        _Pragma("unroll") for (int v = 0; v < GS; ++v) {                        \
            if (((MASK >> (u * GS + v)) & 1) && (COND)) { BODY }                \
        }                                                                       \
    }
#define MB_OFFDIAG (!(DIAG && u == v))
    if (!COARSE) {                                   // level 0: Phi^T Phi (ps holds phi(f)), symmetric
        MB_FOR_BLOCKS(true, dmma(am[u][v][0], am[u][v][1], sr[u], scol[v]);)
        if (MODE == 1) { MB_FOR_BLOCKS(true, dmma(av[u][v][0], av[u][v][1], sr[u] * sr[u], scol[v] * scol[v]);) }
    } else if (MODE == 2) {
        MB_FOR_BLOCKS(true, dmma(am[u][v][0], am[u][v][1], dr[u], dcol[v]);)
    } else {
        MB_FOR_BLOCKS(true, dmma(am[u][v][0], am[u][v][1], sr[u], dcol[v]);)
        MB_FOR_BLOCKS(MB_OFFDIAG, dmma(am[u][v][0], am[u][v][1], dr[u], scol[v]);)
        if (MODE == 1) {
            double ar[GS], br[GS], er[GS], ac[GS], bc[GS], ec[GS];
#pragma unroll
            for (int u = 0; u < GS; ++u) {
                ar[u] = dr[u] * dr[u];
                br[u] = sr[u] * sr[u];
                er[u] = dr[u] * sr[u];
                ac[u] = DIAG ? ar[u] : dcol[u] * dcol[u];
                bc[u] = DIAG ? br[u] : scol[u] * scol[u];
                ec[u] = DIAG ? er[u] : dcol[u] * scol[u];
            }
            MB_FOR_BLOCKS(true, dmma(av[u][v][0], av[u][v][1], ar[u], bc[v]);)
            MB_FOR_BLOCKS(MB_OFFDIAG, dmma(av[u][v][0], av[u][v][1], br[u], ac[v]);)
            MB_FOR_BLOCKS(true, dmma(av[u][v][0], av[u][v][1], MB_OFFDIAG ? er[u] + er[u] : er[u], ec[v]);)
        }
    }
#undef MB_OFFDIAG
#undef MB_FOR_BLOCKS
}

// All k-steps of one tile for one task slot.  The fragment pointers of the slot are computed once per tile; inside
// the 4-k-step unrolled body every shared-memory address is pointer + immediate (constant LD, the row skew has
// period 4 k-steps), so the loop carries no address arithmetic and the loads of a k-step can be hoisted over the DMMAs
// of the previous one.  NS is a multiple of 16.
template <bool COARSE, int MODE, int GS, int MASK, bool DIAG, int LD>
__device__ __forceinline__ void slot_tile(double (&am)[GS][GS][2], double (&av)[GS][GS][2], const double* phi_f,
                                          const double* phi_c, const int (&off_r)[GS], const int (&off_c)[GS],
                                          int NS) {
    for (int k0 = 0; k0 < NS; k0 += 16) {
        const double* pf = phi_f + (size_t)k0 * LD;
        const double* pc = phi_c + (size_t)k0 * LD;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const double* pfq = pf + q * 4 * LD + q;
            const double* pcq = pc + q * 4 * LD + q;
            slot_mma<COARSE, MODE, GS, MASK, DIAG>(am, av, pfq, pcq, off_r, off_c);  // skew of rows k0 + 4q .. : q
        }
    }
}

// the block masks the planner can emit (make_plan).  Off the block diagonal: whole group, one column, one row (edge
// groups), the lower block of the first column, single block; on it: upper triangle, its first row, its last block,
// single block.
#define MB_SLOT_SWITCH(MASKVAR, DIAGVAR, CALL)                                  \
    if (DIAGVAR) {                                                              \
        switch (MASKVAR) {                                                      \
            case 0xB: CALL(0xB, true); break;                                   \
            case 0x3: CALL(0x3, true); break;                                   \
            case 0x8: CALL(0x8, true); break;                                   \
            case 0x1: CALL(0x1, true); break;                                   \
            default: break;                                                     \
        }                                                                       \
    } else {                                                                    \
        switch (MASKVAR) {                                                      \
            case 0xF: CALL(0xF, false); break;                                  \
            case 0x5: CALL(0x5, false); break;                                  \
            case 0xA: CALL(0xA, false); break;                                  \
            case 0x3: CALL(0x3, false); break;                                  \
            case 0x4: CALL(0x4, false); break;                                  \
            case 0x1: CALL(0x1, false); break;                                  \
            default: break;                                                     \
        }                                                                       \
    }

// MODE 0: covariance sums only; 1: covariance sums + sums of squares; 2: Gram of the differences
//
// A tile is PRODUCED (basis rows, one (sample, side) recurrence per thread on the FP64 pipe), then CONSUMED by all 16
// warps with DMMA, in separate phases.  A warp-specialised variant that overlapped the two (4 producer warps filling
// a second tile while 12 warps contracted the first) was slower: an FP64-pipe instruction issued while DMMAs are in
// flight costs the tensor pipe ~10 cycles (measured: contraction alone 33.5 TFLOP/s, producers alone 1/3 of that time,
// together 26.6 TFLOP/s), so the phases do not overlap for free and the larger single tile wins (28.6 TFLOP/s).
// WARPS / SLOTS: 16 warps x 2 task slots, one CTA per SM.  (Two 8-warp CTAs per SM on half-size tiles, so that one
// CTA's produce phase overlaps the other's DMMA stream, lost to this with the S / D tiles and was removed.)
template <bool COARSE, int MODE, int GS, int WARPS, int SLOTS, int LD>
__global__ void __launch_bounds__(WARPS * 32, WARPS == 16 ? 1 : 2) gram_kernel(const GramArgs a) {
    extern __shared__ double sm[];
    const GramPlan& pl = a.plan;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int NS = pl.ns, nb = pl.nb, r_pad = 8 * nb, R = a.basis.size;
    // arrays of NS rows in the tile: S and D (fine + coarse covariance); D alone (difference Gram); phi(f) (level 0)
    constexpr bool ONLY_D = COARSE && MODE == 2;
    constexpr int N_ARRAYS = (COARSE && !ONLY_D) ? 2 : 1;
    const size_t tile_elems = (size_t)N_ARRAYS * NS * LD;
    __shared__ unsigned cnt_sm[2];
    if (tid == 0) cnt_sm[0] = cnt_sm[1] = 0;

    const int my_tasks = pl.n_tasks[warp];
    double acc_m[SLOTS][GS][GS][2];
    double acc_v[MODE == 1 ? SLOTS : 1][GS][GS][2];
#pragma unroll
    for (int s = 0; s < SLOTS; ++s)
#pragma unroll
        for (int u = 0; u < GS; ++u)
#pragma unroll
            for (int v = 0; v < GS; ++v) {
                acc_m[s][u][v][0] = acc_m[s][u][v][1] = 0.0;
                if (MODE == 1) acc_v[s][u][v][0] = acc_v[s][u][v][1] = 0.0;
            }

    const int64_t n_tiles = (a.n + NS - 1) / NS;
    const int frag_off = (lane & 3) * LD + (lane >> 2);
    unsigned cnt_ok = 0, cnt_rm = 0;

    // decode this warp's task list ONCE: shared-memory column offsets of its row / column fragments and the mask of
    // 8x8 blocks each slot owns (0 = empty slot)
    int off_r[SLOTS][GS], off_c[SLOTS][GS], mask[SLOTS];
    bool diag[SLOTS];
#pragma unroll
    for (int slot = 0; slot < SLOTS; ++slot) {
        const bool used = slot < my_tasks;
        diag[slot] = used && pl.gi[warp][slot] == pl.gj[warp][slot];
        const int bi0 = used ? pl.gi[warp][slot] * GS : 0, bj0 = used ? pl.gj[warp][slot] * GS : 0;
#pragma unroll
        for (int u = 0; u < GS; ++u) {
            off_r[slot][u] = 8 * min(bi0 + u, nb - 1);
            off_c[slot][u] = 8 * min(bj0 + u, nb - 1);
        }
        mask[slot] = used ? pl.mk[warp][slot] : 0;
    }

    const double qnan = __longlong_as_double(0x7ff8000000000000LL);
    // Basis rows of one tile: item = sample (fine + coarse levels: both recurrences, rows S and D), at most one item per
    // producing thread.  The raw values of an item are FETCHED one tile ahead (registers), so the DRAM latency hides
    // behind the contraction of the current tile.
    // A sample the mask of a vector quantity drops is fetched as NaN: the domain test below then drops it here too.
    // Legendre: item = (sample, parity of the moments) whenever the CTA has two threads per sample of the tile -- threads
    // [0, NS) take the even moments, [NS, 2 NS) the odd ones (write_rows_parity).
    const double* const pairs = a.pairs + (int64_t)blockIdx.y * a.stride_m;
    const bool split = a.basis.kind == MLMCB200_LEGENDRE && 2 * NS <= WARPS * 32 && (NS & 31) == 0;
    const int n_items = split ? 2 * NS : NS;
    // warp index as a value the compiler knows to be warp-uniform (shuffle from lane 0): the branches on it below are
    // uniform, the code behind them runs converged
    const int warp_u = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int parity_u = (split && warp_u >= (NS >> 5)) ? 1 : 0;
    auto fetch = [&](int64_t tile, int item, double& xf, double& xc) {
        xf = xc = qnan;
        if (item < n_items && tile < n_tiles) {
            const int64_t n = tile * NS + (item >= NS ? item - NS : item);
            if (n < a.n && (a.valid == nullptr || a.valid[n])) {
                xf = __ldcs(pairs + n * a.stride_n);
                if (COARSE) xc = __ldcs(pairs + n * a.stride_n + a.stride_side);
            }
        }
    };
    auto generate = [&](int64_t tile, double* tile_base, int item, double xf, double xc) {
        if (split) {
            if (warp_u >= (NS >> 4)) return;                 // uniform: 2 NS / 32 producing warps
        } else if (item >= NS) {
            return;
        }
        const int s = item - parity_u * NS;
        const int64_t n = tile * NS + s;
        // rows are skewed by (s / 4) % 4 doubles: the 16 lanes of a half-warp then store to 16 distinct 8-byte
        // bank pairs (row stride LD = 4 mod 8 alone gives only 4), and a k-step's 4 rows share one skew
        double* row = tile_base + (size_t)s * LD + ((s >> 2) & 3);
        bool good = n < a.n;
        double tf = 0.0, tc = 0.0;
        if (good) {
            tf = a.basis.kind == MLMCB200_RAW ? xf : map_to_ref(a.basis, xf);
            good = moments_finite(a.basis, tf);
            if (COARSE) {
                tc = a.basis.kind == MLMCB200_RAW ? xc : map_to_ref(a.basis, xc);
                good = good && moments_finite(a.basis, tc);
            }
            if (!parity_u) {
                cnt_ok += good ? 1u : 0u;
                cnt_rm += good ? 0u : 1u;
            }
        }
        double* const row2 = (COARSE && !ONLY_D) ? row + (size_t)NS * LD : row;
        if (split) {
            tf = good ? tf : 0.0;
            tc = good ? tc : 0.0;
            if (parity_u)
                write_rows_parity<COARSE, ONLY_D, 1>(R, tf, tc, good, row, row2, r_pad);
            else
                write_rows_parity<COARSE, ONLY_D, 0>(R, tf, tc, good, row, row2, r_pad);
        } else if (COARSE) {
            write_rows_sd<ONLY_D>(a.basis, tf, tc, good, row, row2, r_pad);
        } else {
            write_row(a.basis, tf, good, row, r_pad);
        }
    };
    auto consume = [&](const double* tile_base) {
        const double* phi_f = tile_base + frag_off;
        const double* phi_c = N_ARRAYS == 2 ? phi_f + (size_t)NS * LD : phi_f;
#pragma unroll
        for (int slot = 0; slot < SLOTS; ++slot) {
            if (GS == 1) {
                if (mask[slot] && diag[slot])
                    slot_tile<COARSE, MODE, GS, 0x1, true, LD>(acc_m[slot], acc_v[MODE == 1 ? slot : 0], phi_f, phi_c,
                                                               off_r[slot], off_c[slot], NS);
                else if (mask[slot])
                    slot_tile<COARSE, MODE, GS, 0x1, false, LD>(acc_m[slot], acc_v[MODE == 1 ? slot : 0], phi_f, phi_c,
                                                                off_r[slot], off_c[slot], NS);
            } else {
#define MB_CALL(M, D) slot_tile<COARSE, MODE, GS, M, D, LD>(acc_m[slot], acc_v[MODE == 1 ? slot : 0], phi_f, phi_c, off_r[slot], off_c[slot], NS)
                MB_SLOT_SWITCH(mask[slot], diag[slot], MB_CALL)   // warp-uniform, once per slot and tile
#undef MB_CALL
            }
        }
    };

    double xf_n, xc_n;                                      // raw values of this thread's item of the next tile
    fetch(blockIdx.x, tid, xf_n, xc_n);
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        generate(tile, sm, tid, xf_n, xc_n);
        fetch(tile + gridDim.x, tid, xf_n, xc_n);
        __syncthreads();
        consume(sm);
        __syncthreads();
    }

    (void)tile_elems;
    // ---- epilogue: one partial [2 + 2 R R] per CTA, upper blocks mirrored ----
    double* const out = a.partial + ((int64_t)blockIdx.y * gridDim.x + blockIdx.x) * a.partial_stride;
    double* const out_m = out + 2;
    double* const out_v = out + 2 + (int64_t)R * R;
#pragma unroll
    for (int slot = 0; slot < SLOTS; ++slot) {
        if (slot < my_tasks) {
            const int bi0 = pl.gi[warp][slot] * GS, bj0 = pl.gj[warp][slot] * GS;
#pragma unroll
            for (int u = 0; u < GS; ++u)
#pragma unroll
                for (int v = 0; v < GS; ++v) {
                    const int I = bi0 + u, J = bj0 + v;
                    if ((pl.mk[warp][slot] >> (u * GS + v)) & 1) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int i = 8 * I + (lane >> 2), j = 8 * J + 2 * (lane & 3) + e;
                            if (i >= R || j >= R) continue;
                            // Legendre tiles hold the monic W_i: P_i P_j = (g_i g_j) W_i W_j
                            const double gg = a.basis.kind == MLMCB200_LEGENDRE ? kLegAlpha[i] * kLegAlpha[j] : 1.0;
                            if (COARSE && MODE != 2) {
                                // accumulators hold 2 sum d_ij (4 sum d_ij^2); a DIAGONAL block holds the one-sided
                                // X = S_I^T D_I (Y = A_I^T B_I + E_I^T E_I) and is written un-mirrored: the reduction
                                // (reduce_partials_sym_kernel) forms (T + T^T) / 2 of the whole matrix
                                const double hm = I == J ? 1.0 : 0.5, hv = I == J ? 0.5 : 0.25;
                                const double vm = acc_m[slot][u][v][e] * gg * hm;
                                out_m[(int64_t)i * R + j] = vm;
                                if (I != J) out_m[(int64_t)j * R + i] = vm;
                                if (MODE == 1) {
                                    const double vv = acc_v[slot][u][v][e] * (gg * gg) * hv;
                                    out_v[(int64_t)i * R + j] = vv;
                                    if (I != J) out_v[(int64_t)j * R + i] = vv;
                                }
                            } else if (I != J || j >= i) {
                                // symmetric products: diagonal blocks keep the upper triangle and mirror it
                                const double vm = acc_m[slot][u][v][e] * gg;
                                out_m[(int64_t)i * R + j] = vm;
                                if (i != j) out_m[(int64_t)j * R + i] = vm;
                                if (MODE == 1) {
                                    const double vv = acc_v[slot][u][v][e] * (gg * gg);
                                    out_v[(int64_t)i * R + j] = vv;
                                    if (i != j) out_v[(int64_t)j * R + i] = vv;
                                }
                            }
                        }
                    }
                }
        }
    }
    if (cnt_ok) atomicAdd(&cnt_sm[0], cnt_ok);
    if (cnt_rm) atomicAdd(&cnt_sm[1], cnt_rm);
    __syncthreads();
    if (tid == 0) {
        out[0] = (double)cnt_sm[0];
        out[1] = (double)cnt_sm[1];
    }
}

// acc[0..1] += counts; acc[2 + k] += (P[k] + P[k^T]) / 2 with P = sum of the partials (fixed order) and k^T the
// transposed index inside each R x R matrix (sums, then sums of squares).  Entries that the kernel wrote mirrored are
// bit-identical on both sides ((a + a) / 2 = a); the one-sided diagonal blocks of the fine + coarse modes become X + X^T.
// One warp per output: lanes stride over the partials, shuffle tree.
// Vector quantities: blockIdx.y = component of this launch (global number comp0 + blockIdx.y of n_comp); its matrices go
// to acc[2 + mat * n_comp R^2 + comp R^2 + e] (flat index m R R + i R + j of quantity_estimate.py:131-147), the counts
// (equal for all components: they share the sample mask) are added once, by component 0.
__global__ void reduce_partials_sym_kernel(const double* __restrict__ partial, int n_partials, int64_t stride, int R,
                                           int n_mats, double* __restrict__ acc, int comp0, int n_comp) {
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t R2 = (int64_t)R * R, len = 2 + n_mats * R2;
    if (j >= len) return;
    const int comp = comp0 + blockIdx.y;
    partial += (int64_t)blockIdx.y * n_partials * stride;
    int64_t jt = j, jo = j;
    if (j >= 2) {
        const int64_t k = j - 2, mat = k / R2, e = k - mat * R2;
        const int64_t row = e / R, col = e - row * R;
        jt = 2 + mat * R2 + col * R + row;
        jo = 2 + (mat * n_comp + comp) * R2 + e;
    } else if (comp != 0) {
        return;
    }
    double s1 = 0.0, s2 = 0.0;
    for (int b0 = lane; b0 < n_partials; b0 += 32 * 5) {      // batches of independent loads, added in index order
        double v1[5], v2[5];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            const int b = b0 + 32 * u;
            v1[u] = b < n_partials ? partial[(int64_t)b * stride + j] : 0.0;
            v2[u] = b < n_partials ? partial[(int64_t)b * stride + jt] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            s1 += v1[u];
            s2 += v2[u];
        }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (lane == 0) acc[jo] += j < 2 ? s1 : 0.5 * (s1 + s2);
}

// ---- task planner ------------------------------------------------------------------------------------------------
// The upper triangle of 8x8 blocks is cut into tasks (a GS x GS group of blocks, or a part of one that the kernel has a
// straight-line specialisation for) and dealt to the warps.  What has to be balanced is the DMMA count per SM
// SUB-PARTITION (warp w issues on sub-partition w % 4; each has its own FP64 tensor pipe, and two resident warps keep it
// busy), with the cost of a block depending on the mode: a diagonal block needs one product where an off-diagonal one
// needs two (sums of fine + coarse levels), 3 vs 5 with the sums of squares.  The planner tries every small number of
// splits of each kind, packs the pieces into the four sub-partitions (longest processing time first + pairwise
// improvement) and keeps the split with the lowest maximum; inside a sub-partition the pieces go to its warps the same way.
struct PlanTask { int gi, gj, mask, cost; };

int block_cost_sum(int gi, int gj, int mask, int gs, int cd, int co) {
    int c = 0;
    for (int u = 0; u < gs; ++u)
        for (int v = 0; v < gs; ++v)
            if ((mask >> (u * gs + v)) & 1) c += (gi * gs + u == gj * gs + v) ? cd : co;
    return c;
}

// pieces -> n_bins bins of at most cap pieces; returns the largest bin load (or -1 if they do not fit)
int pack_bins(const PlanTask* pieces, int n, int n_bins, int cap, int* bin_of) {
    int order[kWarps * kMaxSlots];
    for (int i = 0; i < n; ++i) order[i] = i;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j)
            if (pieces[order[j]].cost > pieces[order[i]].cost) { const int t = order[i]; order[i] = order[j]; order[j] = t; }
    int load[kWarps] = {0}, cnt[kWarps] = {0};
    for (int k = 0; k < n; ++k) {
        int best = -1;
        for (int b = 0; b < n_bins; ++b)
            if (cnt[b] < cap && (best < 0 || load[b] < load[best])) best = b;
        if (best < 0) return -1;
        bin_of[order[k]] = best;
        load[best] += pieces[order[k]].cost;
        ++cnt[best];
    }
    // improvement: move or swap pieces between the heaviest bin and another one while that lowers the pair's maximum
    for (int pass = 0; pass < 64; ++pass) {
        int hb = 0;
        for (int b = 1; b < n_bins; ++b) if (load[b] > load[hb]) hb = b;
        bool changed = false;
        for (int i = 0; i < n && !changed; ++i) {
            if (bin_of[i] != hb) continue;
            for (int ob = 0; ob < n_bins && !changed; ++ob) {
                if (ob == hb) continue;
                if (cnt[ob] < cap && load[ob] + pieces[i].cost < load[hb]) {            // move
                    bin_of[i] = ob; load[hb] -= pieces[i].cost; load[ob] += pieces[i].cost; --cnt[hb]; ++cnt[ob];
                    changed = true;
                    break;
                }
                for (int j = 0; j < n; ++j) {                                           // swap
                    if (bin_of[j] != ob) continue;
                    const int d = pieces[i].cost - pieces[j].cost;
                    if (d > 0 && load[ob] + d < load[hb]) {
                        bin_of[i] = ob; bin_of[j] = hb; load[hb] -= d; load[ob] += d;
                        changed = true;
                        break;
                    }
                }
            }
        }
        if (!changed) break;
    }
    int worst = 0;
    for (int b = 0; b < n_bins; ++b) worst = load[b] > worst ? load[b] : worst;
    return worst;
}

// n_warps / n_slots: warps that contract and task slots per warp; cd / co: DMMAs per diagonal / off-diagonal block
int make_plan_uncached(int R, int n_warps, int n_slots, GramPlan* pl, size_t* smem, int fixed_ld, int ns_cap,
                       size_t budget, int cd, int co, int n_arrays) {
    const int nb = (R + 7) / 8;
    if (fixed_ld && R > kGramMaxMoments) {
        set_error("gram: %d moments do not fit the shared-memory tile (max %d)", R, kGramMaxMoments);
        return -1;
    }
    pl->nb = nb;
    pl->gs = nb <= 5 ? 1 : 2;
    const int gs = pl->gs;
    const int ng = (nb + gs - 1) / gs;
    const int kMaxTasks = n_warps * n_slots;
    // natural tasks: one per pair of block groups on/above the diagonal; mask = blocks that exist and have J >= I
    PlanTask nat[kWarps * kMaxSlots];
    int n_nat = 0, total = 0;
    for (int gi = 0; gi < ng; ++gi)
        for (int gj = gi; gj < ng; ++gj) {
            int mask = 0;
            for (int u = 0; u < gs; ++u)
                for (int v = 0; v < gs; ++v) {
                    const int I = gi * gs + u, J = gj * gs + v;
                    if (I < nb && J < nb && J >= I) mask |= 1 << (u * gs + v);
                }
            if (n_nat >= kMaxTasks) {
                set_error("gram: %d moments need more than %d block tasks", R, kMaxTasks);
                return -1;
            }
            nat[n_nat] = {gi, gj, mask, block_cost_sum(gi, gj, mask, gs, cd, co)};
            total += nat[n_nat++].cost;
        }
    // splits the kernel has specialisations for: whole off-diagonal group -> two columns (0xF -> 0x5 + 0xA), diagonal
    // triangle -> first row + last block (0xB -> 0x3 + 0x8), off-diagonal column -> two blocks (0x5 -> 0x1 + 0x4)
    int n_full = 0, n_tri = 0, n_col = 0;
    for (int i = 0; i < n_nat; ++i) {
        const bool dg = nat[i].gi == nat[i].gj;
        n_full += !dg && nat[i].mask == 0xF;
        n_tri += dg && nat[i].mask == 0xB;
        n_col += !dg && nat[i].mask == 0x5;
    }
    const int n_bins = n_warps < 4 ? n_warps : 4, cap = (n_warps / n_bins) * n_slots;
    const int ideal = (total + n_bins - 1) / n_bins;
    PlanTask best[kWarps * kMaxSlots];
    int best_bin[kWarps * kMaxSlots], best_n = 0, best_worst = -1, best_warp = 0;
    (void)ideal;
    for (int sa = 0; sa <= (gs == 2 ? n_full : 0) && sa <= 8; ++sa)
        for (int sb = 0; sb <= (gs == 2 ? n_tri : 0) && sb <= 6; ++sb)
            for (int sc = 0; sc <= (gs == 2 ? n_col : 0) && sc <= 6; ++sc) {
                if (n_nat + sa + sb + sc > kMaxTasks) continue;
                PlanTask pc[kWarps * kMaxSlots];
                int n = 0, ua = 0, ub = 0, uc = 0;
                for (int i = 0; i < n_nat; ++i) {
                    const PlanTask t = nat[i];
                    const bool dg = t.gi == t.gj;
                    int m1 = t.mask, m2 = 0;
                    if (!dg && t.mask == 0xF && ua < sa) { m1 = 0x5; m2 = 0xA; ++ua; }
                    else if (dg && t.mask == 0xB && ub < sb) { m1 = 0x3; m2 = 0x8; ++ub; }
                    else if (!dg && t.mask == 0x5 && uc < sc) { m1 = 0x1; m2 = 0x4; ++uc; }
                    pc[n++] = {t.gi, t.gj, m1, block_cost_sum(t.gi, t.gj, m1, gs, cd, co)};
                    if (m2) pc[n++] = {t.gi, t.gj, m2, block_cost_sum(t.gi, t.gj, m2, gs, cd, co)};
                }
                int bin_of[kWarps * kMaxSlots];
                const int worst = pack_bins(pc, n, n_bins, cap, bin_of);
                if (worst < 0) continue;
                // second criterion: the heaviest WARP (a sub-partition whose load sits in one or two warps leaves the
                // pipe to a lone warp for part of the tile, and a lone warp does not hide its own latencies)
                int warp_worst = 0;
                bool fits = true;
                for (int b = 0; b < n_bins && fits; ++b) {
                    PlanTask mine[kWarps * kMaxSlots];
                    int m = 0, warp_of[kWarps * kMaxSlots];
                    for (int i = 0; i < n; ++i)
                        if (bin_of[i] == b) mine[m++] = pc[i];
                    const int ww = pack_bins(mine, m, n_warps / n_bins, n_slots, warp_of);
                    fits = ww >= 0;
                    warp_worst = ww > warp_worst ? ww : warp_worst;
                }
                if (!fits) continue;
                if (best_worst < 0 || worst < best_worst || (worst == best_worst && warp_worst < best_warp) ||
                    (worst == best_worst && warp_worst == best_warp && n < best_n)) {
                    best_worst = worst;
                    best_warp = warp_worst;
                    best_n = n;
                    for (int i = 0; i < n; ++i) { best[i] = pc[i]; best_bin[i] = bin_of[i]; }
                }
            }
    if (best_worst < 0) {
        set_error("gram: no task plan for %d moments", R);
        return -1;
    }
    // inside a sub-partition: its pieces to its warps (b, b + 4, ...), at most n_slots each
    for (int w = 0; w < kWarps; ++w) pl->n_tasks[w] = 0;
    for (int b = 0; b < n_bins; ++b) {
        PlanTask mine[kWarps * kMaxSlots];
        int n = 0;
        for (int i = 0; i < best_n; ++i)
            if (best_bin[i] == b) mine[n++] = best[i];
        int warp_of[kWarps * kMaxSlots];
        const int n_sub = n_warps / n_bins;
        if (pack_bins(mine, n, n_sub, n_slots, warp_of) < 0) {
            set_error("gram: task plan for %d moments does not fit the warps", R);
            return -1;
        }
        for (int i = 0; i < n; ++i) {
            const int w = b + n_bins * warp_of[i];
            const int k = pl->n_tasks[w]++;
            pl->gi[w][k] = (unsigned char)mine[i].gi;
            pl->gj[w][k] = (unsigned char)mine[i].gj;
            pl->mk[w][k] = (unsigned char)mine[i].mask;
        }
    }
    int ld = 8 * nb;
    while (ld % 8 != 4) ++ld;
    if (fixed_ld) ld = fixed_ld;
    pl->ld = ld;
    // tile of n_arrays arrays of NS rows; NS a multiple of 4
    int ns = (int)(budget / ((size_t)n_arrays * ld * sizeof(double)));
    if (ns > ns_cap) ns = ns_cap;
    ns = fixed_ld ? (ns / 16) * 16 : (ns / 4) * 4;      // covariance tiles: whole 4-k-step unrolled bodies
    if (ns < 8) {
        set_error("gram: %d moments do not fit the shared-memory tile", R);
        return -1;
    }
    pl->ns = ns;
    *smem = (size_t)n_arrays * ns * ld * sizeof(double);
    return 0;
}

// The search above costs ~1 ms of host time: plans are memoised per thread (a handful of distinct keys per process).
int make_plan(int R, int n_warps, int n_slots, GramPlan* pl, size_t* smem, int fixed_ld = 0, int ns_cap = 128,
              size_t budget = 216u * 1024u, int cd = 1, int co = 1, int n_arrays = 2) {
    constexpr int kKey = 9;
    struct Entry { int key[kKey]; GramPlan plan; size_t smem; };
    constexpr int kCache = 16;
    static thread_local Entry cache[kCache];
    static thread_local int n_cached = 0, next = 0;
    const int key[kKey] = {R, n_warps, n_slots, fixed_ld, ns_cap, (int)budget, cd, co, n_arrays};
    for (int i = 0; i < n_cached; ++i) {
        bool same = true;
        for (int k = 0; k < kKey; ++k) same = same && cache[i].key[k] == key[k];
        if (same) {
            *pl = cache[i].plan;
            *smem = cache[i].smem;
            return 0;
        }
    }
    const int rc = make_plan_uncached(R, n_warps, n_slots, pl, smem, fixed_ld, ns_cap, budget, cd, co, n_arrays);
    if (rc != 0) return rc;
    Entry& e = cache[next];
    for (int k = 0; k < kKey; ++k) e.key[k] = key[k];
    e.plan = *pl;
    e.smem = *smem;
    next = (next + 1) % kCache;
    if (n_cached < kCache) ++n_cached;
    return 0;
}

template <bool COARSE, int MODE, int GS, int LD>
int launch_gram(const GramArgs& a, int grid, int n_comp, size_t smem, cudaStream_t st) {
    auto kern = gram_kernel<COARSE, MODE, GS, kWarps, 2, LD>;
    MB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)grid, (unsigned)n_comp), kWarps * 32, smem, st>>>(a);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}

// tile width and block-group size follow from the number of 8-moment blocks (gram_ld, make_plan): up to 4 blocks
// (narrow tile, single blocks), 5 (mid tile, single blocks), 6-7 (mid tile, 2 x 2 groups), 8-13 (wide tile, 2 x 2 groups)
template <bool COARSE, int MODE>
int launch_gram_gs(const GramArgs& a, int grid, int n_comp, size_t smem, cudaStream_t st) {
    const int nb = a.plan.nb;
    if (nb <= 4) return launch_gram<COARSE, MODE, 1, kLDSmall>(a, grid, n_comp, smem, st);
    if (nb <= 5) return launch_gram<COARSE, MODE, 1, kLDMid>(a, grid, n_comp, smem, st);
    if (nb <= 7) return launch_gram<COARSE, MODE, 2, kLDMid>(a, grid, n_comp, smem, st);
    return launch_gram<COARSE, MODE, 2, kLDWide>(a, grid, n_comp, smem, st);
}

// ------------------------------------------------------------------------------------------------------------
// Max-entropy functional pieces on a fixed node set (mlmc/tool/simple_distribution.py:254-327):
//   rho_q = exp(clip(-Phi_q . lam, -200, 200));  F = sum w rho;  g_i = sum w rho Phi_qi;  H = Phi^T diag(w rho) Phi
// Same tiling as the covariance kernel: a CTA stages a tile of Phi rows in shared memory (coalesced copy), one
// warp per node forms the exponent, and H goes through DMMA with A = (w rho) . Phi^T and B = Phi.
// ------------------------------------------------------------------------------------------------------------
struct MaxentArgs {
    const double* phi;
    int64_t ld_g;           // row stride of phi in global memory
    const double* w;
    const double* lam;
    int64_t n_nodes;
    int R;
    int want_h;
    double* partial;        // [gridDim.x][1 + R + R R]
    int64_t partial_stride;
    GramPlan plan;
};

template <int GS>
__global__ void __launch_bounds__(kThreadsGram, 1) maxent_kernel(const MaxentArgs a) {
    extern __shared__ double sm[];
    const GramPlan& pl = a.plan;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int LD = pl.ld, NS = pl.ns, nb = pl.nb, r_pad = 8 * nb, R = a.R;
    double* const phi_s = sm;                                  // [NS][LD]
    double* const wr = sm + (size_t)NS * LD;                   // [NS]  w_q rho_q
    double* const lam_s = wr + NS;                             // [r_pad]
    for (int i = tid; i < r_pad; i += kThreadsGram) lam_s[i] = i < R ? a.lam[i] : 0.0;

    const int my_tasks = pl.n_tasks[warp];
    double acc[kMaxSlots][GS][GS][2];
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s)
#pragma unroll
        for (int u = 0; u < GS; ++u)
#pragma unroll
            for (int v = 0; v < GS; ++v) acc[s][u][v][0] = acc[s][u][v][1] = 0.0;
    // F and g: the warp that forms the exponent of a node also adds the node into its partial sums
    // (lane l owns moments l, l + 32, ...); the per-warp partials meet in shared memory at the end
    constexpr int GJ = MLMCB200_MAX_MOMENTS / 32;
    double g_loc[GJ], f_loc = 0.0;
#pragma unroll
    for (int j = 0; j < GJ; ++j) g_loc[j] = 0.0;
    const int frag_off = (lane & 3) * LD + (lane >> 2);
    const int64_t n_tiles = (a.n_nodes + NS - 1) / NS;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t q0 = tile * NS;
        __syncthreads();
        // one warp per node: stage the row (coalesced, zero padded), exponent, weight, F / g partials
        for (int s = warp; s < NS; s += kWarps) {
            const int64_t q = q0 + s;
            const bool in = q < a.n_nodes;
            double* row = phi_s + (size_t)s * LD;
            double dot = 0.0;
            double mine[GJ];
#pragma unroll
            for (int j = 0; j < GJ; ++j) {
                const int i = lane + 32 * j;
                mine[j] = 0.0;
                if (i < r_pad) {
                    mine[j] = (in && i < R) ? __ldg(a.phi + q * a.ld_g + i) : 0.0;
                    row[i] = mine[j];
                    dot = fma(mine[j], lam_s[i], dot);
                }
            }
            dot = warp_sum(dot);
            const double power = fmin(fmax(-dot, -200.0), 200.0);
            const double wq = in ? __ldg(a.w + q) * exp(power) : 0.0;
            if (lane == 0) wr[s] = wq;
            f_loc += wq;
#pragma unroll
            for (int j = 0; j < GJ; ++j) g_loc[j] = fma(wq, mine[j], g_loc[j]);
        }
        __syncthreads();
        if (a.want_h) {
            for (int k0 = 0; k0 < NS; k0 += 4) {
                const double* pf = phi_s + (size_t)k0 * LD + frag_off;
                const double wq = wr[k0 + (lane & 3)];
#pragma unroll
                for (int slot = 0; slot < kMaxSlots; ++slot) {
                    if (slot < my_tasks) {
                        const int bi0 = pl.gi[warp][slot] * GS, bj0 = pl.gj[warp][slot] * GS;
                        double fr[GS], fcol[GS];
#pragma unroll
                        for (int u = 0; u < GS; ++u) {
                            fr[u] = bi0 + u < nb ? pf[8 * (bi0 + u)] * wq : 0.0;
                            fcol[u] = bj0 + u < nb ? pf[8 * (bj0 + u)] : 0.0;
                        }
#pragma unroll
                        for (int u = 0; u < GS; ++u)
#pragma unroll
                            for (int v = 0; v < GS; ++v) {
                                const int I = bi0 + u, J = bj0 + v;
                                (void)I; (void)J;
                                if ((pl.mk[warp][slot] >> (u * GS + v)) & 1) dmma(acc[slot][u][v][0], acc[slot][u][v][1], fr[u], fcol[v]);
                            }
                    }
                }
            }
        }
    }

    double* const out = a.partial + (int64_t)blockIdx.x * a.partial_stride;
    // per-warp F / g partials -> shared memory (the tile buffer is free now) -> one value per moment
    __syncthreads();
    double* const red = sm;                                     // [kWarps][1 + 32 * GJ]
    constexpr int kRedLd = 1 + 32 * GJ;
#pragma unroll
    for (int j = 0; j < GJ; ++j) red[warp * kRedLd + 1 + lane + 32 * j] = g_loc[j];
    if (lane == 0) red[warp * kRedLd] = f_loc;
    __syncthreads();
    for (int i = tid; i < 1 + R; i += kThreadsGram) {
        double v = 0.0;
        for (int w = 0; w < kWarps; ++w) v += red[w * kRedLd + i];
        out[i] = v;
    }
    __syncthreads();
    if (a.want_h) {
        double* const out_h = out + 1 + R;
#pragma unroll
        for (int slot = 0; slot < kMaxSlots; ++slot) {
            if (slot < my_tasks) {
                const int bi0 = pl.gi[warp][slot] * GS, bj0 = pl.gj[warp][slot] * GS;
#pragma unroll
                for (int u = 0; u < GS; ++u)
#pragma unroll
                    for (int v = 0; v < GS; ++v) {
                        const int I = bi0 + u, J = bj0 + v;
                        if ((pl.mk[warp][slot] >> (u * GS + v)) & 1) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) {
                                const int i = 8 * I + (lane >> 2), j = 8 * J + 2 * (lane & 3) + e;
                                if (i < R && j < R && (I != J || j >= i)) {
                                    out_h[(int64_t)i * R + j] = acc[slot][u][v][e];
                                    if (i != j) out_h[(int64_t)j * R + i] = acc[slot][u][v][e];
                                }
                            }
                        }
                    }
            }
        }
    }
}

// out[j] = sum_b partial[b][j]: one warp per output, lanes stride over the partials, shuffle tree (fixed order)
__global__ void sum_partials_kernel(const double* __restrict__ partial, int n_partials, int64_t stride, int64_t len,
                                    double* __restrict__ out) {
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (j >= len) return;
    double s = 0.0;
    for (int b = lane; b < n_partials; b += 32) s += partial[(int64_t)b * stride + j];
    s = warp_sum(s);
    if (lane == 0) out[j] = s;
}

int maxent_grid(int64_t n_nodes, int ns) {
    const int64_t tiles = (n_nodes + ns - 1) / ns;
    int grid = sm_count();
    if (tiles < grid) grid = (int)(tiles > 0 ? tiles : 1);
    return grid;
}
}  // namespace
}  // namespace mlmcb200

using namespace mlmcb200;

extern "C" int64_t mlmcb200_gram_workspace_bytes(int32_t size) {
    return mlmcb200_gram_workspace_bytes_comp(size, 1);
}

// partials of up to 2 CTAs per SM (scalar quantity) + one per component of a launch (at most kGramCompBatch at a time)
constexpr int kGramCompBatch = 2048;

extern "C" int64_t mlmcb200_gram_workspace_bytes_comp(int32_t size, int32_t n_comp) {
    if (size < 1 || size > MLMCB200_MAX_MOMENTS || n_comp < 1) return -1;
    const int64_t n_partials = 2 * (int64_t)sm_count() + (n_comp > 1 ? (n_comp < kGramCompBatch ? n_comp : kGramCompBatch) : 0);
    return n_partials * (2 + 2 * (int64_t)size * size) * (int64_t)sizeof(double);
}

extern "C" int mlmcb200_gram_accumulate(const mlmcb200_basis_t* basis, const double* pairs, int64_t n,
                                        int64_t stride_n, int64_t stride_side, int32_t has_coarse,
                                        int32_t mode, int32_t want_var, double* acc,
                                        void* workspace, int64_t workspace_bytes, void* stream) {
    return mlmcb200_gram_accumulate_comp(basis, pairs, n, 1, stride_n, stride_side, 0, has_coarse, nullptr, mode,
                                         want_var, acc, workspace, workspace_bytes, stream);
}

extern "C" int mlmcb200_gram_accumulate_comp(const mlmcb200_basis_t* basis, const double* pairs, int64_t n,
                                             int32_t n_comp, int64_t stride_n, int64_t stride_side, int64_t stride_m,
                                             int32_t has_coarse, const uint8_t* valid, int32_t mode, int32_t want_var,
                                             double* acc, void* workspace, int64_t workspace_bytes, void* stream) {
    if (check_basis(basis) != 0) return -1;
    MB_REQUIRE(n >= 0 && (mode == 0 || mode == 1), "gram_accumulate: bad n=%lld mode=%d", (long long)n, mode);
    MB_REQUIRE(n_comp >= 1, "gram_accumulate: bad n_comp=%d", n_comp);
    MB_REQUIRE(acc != nullptr && workspace != nullptr, "gram_accumulate: null acc/workspace");
    MB_REQUIRE(mode == 0 || has_coarse, "gram_accumulate: mode 1 (difference Gram) needs a coarse side");
    MB_REQUIRE(n_comp == 1 || valid != nullptr,
               "gram_accumulate: a vector quantity needs the sample mask of mlmcb200_sample_mask");
    if (n == 0) return 0;
    MB_REQUIRE(pairs != nullptr, "gram_accumulate: null pairs");
    GramArgs a;
    a.basis = *basis;
    a.n = n;
    a.stride_n = stride_n;
    a.stride_side = stride_side;
    a.stride_m = stride_m;
    a.valid = valid;
    size_t smem = 0;
    // DMMAs per diagonal / off-diagonal block of the mode (slot_mma)
    int cd = 1, co = 1;
    if (mode == 0 && has_coarse) {
        cd = want_var ? 3 : 1;
        co = want_var ? 5 : 2;
    }
    // the tile holds S and D (fine + coarse covariance) or ONE array (level 0: phi(f); difference Gram: D): a single
    // array takes twice the samples.  One sample per thread and tile at most.
    const int n_arrays = (has_coarse && mode == 0) ? 2 : 1;
    const int nb_blocks = (basis->size + 7) / 8;
    if (make_plan(basis->size, kWarps, 2, &a.plan, &smem, gram_ld(nb_blocks), kThreadsGram, 216u * 1024u, cd, co,
                  n_arrays) != 0)
        return -1;
    const int64_t R2 = (int64_t)basis->size * basis->size;
    const int64_t stride = 2 + 2 * R2;
    const int64_t tiles = (n + a.plan.ns - 1) / a.plan.ns;
    // CTAs per component: the SMs are shared out between the components of a launch
    const int sms = sm_count();
    int grid = sms / n_comp;                          // whole waves: n_comp * grid <= SMs (one CTA per SM)
    if (grid < 1) grid = 1;
    if (tiles < grid) grid = (int)tiles;
    const int64_t fit = workspace_bytes / ((int64_t)grid * stride * 8);
    MB_REQUIRE(fit >= 1, "gram_accumulate: workspace too small");
    int batch = n_comp < kGramCompBatch ? n_comp : kGramCompBatch;
    if (fit < batch) batch = (int)fit;
    a.partial = static_cast<double*>(workspace);
    a.partial_stride = stride;
    cudaStream_t st = (cudaStream_t)stream;
    // sums always; sums of squares only when they were produced
    const int n_mats = want_var && mode == 0 ? 2 : 1;
    const int64_t len = 2 + n_mats * R2;
    for (int comp0 = 0; comp0 < n_comp; comp0 += batch) {
        const int nc = n_comp - comp0 < batch ? n_comp - comp0 : batch;
        a.pairs = pairs + (int64_t)comp0 * stride_m;
        int rc;
        if (mode == 1)
            rc = launch_gram_gs<true, 2>(a, grid, nc, smem, st);
        else if (want_var)
            rc = has_coarse ? launch_gram_gs<true, 1>(a, grid, nc, smem, st) : launch_gram_gs<false, 1>(a, grid, nc, smem, st);
        else
            rc = has_coarse ? launch_gram_gs<true, 0>(a, grid, nc, smem, st) : launch_gram_gs<false, 0>(a, grid, nc, smem, st);
        if (rc != 0) return rc;
        reduce_partials_sym_kernel<<<dim3((unsigned)((len * 32 + 255) / 256), (unsigned)nc), 256, 0, st>>>(
            a.partial, grid, stride, basis->size, n_mats, acc, comp0, n_comp);
        MB_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

extern "C" int64_t mlmcb200_maxent_workspace_bytes(int64_t n_nodes, int32_t size) {
    if (size < 1 || size > MLMCB200_MAX_MOMENTS || n_nodes < 0) return -1;
    const int64_t v1 = (int64_t)sm_count() * (1 + (int64_t)size + (int64_t)size * size) * (int64_t)sizeof(double);
    const int64_t v2 = maxent_fast_workspace_bytes(n_nodes, size);
    return v1 > v2 ? v1 : v2;
}

extern "C" int mlmcb200_maxent_fgh(const double* phi, int64_t ld, const double* w, const double* lam_scaled,
                                   int64_t n_nodes, int32_t size, int32_t what, double* out,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
    MB_REQUIRE(phi && w && lam_scaled && out && workspace, "maxent_fgh: null pointer");
    MB_REQUIRE(size >= 1 && size <= MLMCB200_MAX_MOMENTS && ld >= size && n_nodes >= 1, "maxent_fgh: bad sizes");
    MB_REQUIRE((what & 7) != 0, "maxent_fgh: nothing requested");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t stride = 1 + (int64_t)size + (int64_t)size * size;
    const int want_h = (what & 4) ? 1 : 0;
    int grid = 0;
    // up to 104 moments: the kernels of maxent.cu; beyond: the first version below
    int rc = (getenv("MLMCB200_MAXENT_V1") != nullptr) ? 1 : maxent_fast_launch(
        phi, ld, w, lam_scaled, n_nodes, size, want_h, out, workspace, workspace_bytes, st);
    if (rc <= 0) return rc;
    MB_REQUIRE(workspace_bytes >= (int64_t)sm_count() * stride * 8, "maxent_fgh: workspace too small");
    MaxentArgs a;
    a.want_h = want_h;
    a.partial = static_cast<double*>(workspace);
    {
        a.phi = phi;
        a.ld_g = ld;
        a.w = w;
        a.lam = lam_scaled;
        a.n_nodes = n_nodes;
        a.R = size;
        size_t smem = 0;
        if (make_plan(size, kWarps, 2, &a.plan, &smem) != 0) return -1;
        // single table + weights + multipliers instead of two tables + flags
        const int ns = a.plan.ns;
        smem = ((size_t)ns * a.plan.ld + ns + 8 * a.plan.nb) * sizeof(double);
        const size_t red_bytes = (size_t)kWarps * (1 + MLMCB200_MAX_MOMENTS) * sizeof(double);   // F / g partials per warp
        if (smem < red_bytes) smem = red_bytes;
        grid = maxent_grid(n_nodes, ns);
        a.partial_stride = stride;
        auto kern = a.plan.gs == 1 ? maxent_kernel<1> : maxent_kernel<2>;
        MB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kThreadsGram, smem, st>>>(a);
        MB_CUDA_OK(cudaGetLastError());
    }
    const int64_t len = a.want_h ? stride : 1 + size;
    const int threads = 256;
    sum_partials_kernel<<<(unsigned)((len * 32 + threads - 1) / threads), threads, 0, st>>>(a.partial, grid, stride, len,
                                                                                            out);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}
