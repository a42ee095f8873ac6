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
#include "common.cuh"

namespace mlmcb200 {
namespace {

constexpr int kWarps = 8;
constexpr int kThreadsGram = kWarps * 32;
constexpr int kMaxSlots = 4;     // tasks per warp
constexpr int kMaxGroup = 2;     // blocks per task side

struct GramPlan {
    int nb;          // 8x8 blocks per side
    int gs;          // blocks per group (1 or 2)
    int ld;          // smem leading dimension (doubles)
    int ns;          // samples per tile (multiple of 4)
    int n_tasks[kWarps];
    unsigned char gi[kWarps][kMaxSlots];
    unsigned char gj[kWarps][kMaxSlots];
};

struct GramArgs {
    mlmcb200_basis_t basis;
    const double* pairs;
    int64_t n, stride_n, stride_side;
    double* partial;         // [gridDim.x][2 + 2 R R]
    int64_t partial_stride;
    GramPlan plan;
};

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Warp-uniformly predicated DMMA (no branch, no re-convergence barrier around the mma.sync)
__device__ __forceinline__ void dmma_if(unsigned on, double& c0, double& c1, double a, double b) {
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " setp.ne.u32 p, %4, 0;\n"
        " @p mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
        "}\n"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b), "r"(on));
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
        for (int i = 2; i < R; ++i) {
            const double p2 = b.kind == MLMCB200_LEGENDRE ? fma(kLegA[i] * t, p1, -(kLegB[i] * p0)) : p1 * t;
            row[i] = p2;
            p0 = p1;
            p1 = p2;
        }
    }
    for (int i = R; i < r_pad; ++i) row[i] = 0.0;
}

// All DMMAs of one task (GS x GS blocks of shape SHAPE) for one 4-sample step.  Products into the same accumulator
// are issued in separate passes so that consecutive DMMAs are independent.
template <int GS, int SHAPE>
__device__ __forceinline__ constexpr bool block_on(int u, int v) {
    return SHAPE == 0 ? true : SHAPE == 1 ? v == 0 : SHAPE == 2 ? v >= u : (u == 0 && v == 0);
}

template <bool COARSE, int MODE, int GS, int SHAPE>
__device__ __forceinline__ void slot_mma(double (&am)[GS][GS][2], double (&av)[GS][GS][2], const double* pf,
                                         const double* pc, const int (&off_r)[GS], const int (&off_c)[GS]) {
    constexpr int NR = (SHAPE == 3) ? 1 : GS;                 // row / column fragments actually needed
    constexpr int NC = (SHAPE == 1 || SHAPE == 3) ? 1 : GS;
    double fr[GS], cr[GS], fcol[GS], ccol[GS];
#pragma unroll
    for (int u = 0; u < GS; ++u) {
        fr[u] = cr[u] = fcol[u] = ccol[u] = 0.0;
        if (u < NR) {
            fr[u] = pf[off_r[u]];
            if (COARSE) cr[u] = pc[off_r[u]];
        }
    }
#pragma unroll
    for (int v = 0; v < GS; ++v) {
        if (v < NC) {
            if (SHAPE == 2 || SHAPE == 3) {                    // diagonal task: column fragments = row fragments
                fcol[v] = fr[v];
                ccol[v] = cr[v];
            } else {
                fcol[v] = pf[off_c[v]];
                if (COARSE) ccol[v] = pc[off_c[v]];
            }
        }
    }
#define MB_FOR_BLOCKS(BODY)                                                     \
    _Pragma("unroll") for (int u = 0; u < GS; ++u) {                            \
        _Pragma("unroll") for (int v = 0; v < GS; ++v) {                        \
            if (block_on<GS, SHAPE>(u, v)) { BODY }                             \
        }                                                                       \
    }
    if (MODE == 2) {
        MB_FOR_BLOCKS(dmma(am[u][v][0], am[u][v][1], fr[u] - cr[u], fcol[v] - ccol[v]);)
    } else {
        MB_FOR_BLOCKS(dmma(am[u][v][0], am[u][v][1], fr[u], fcol[v]);)
        if (COARSE) { MB_FOR_BLOCKS(dmma(am[u][v][0], am[u][v][1], -cr[u], ccol[v]);) }
        if (MODE == 1) {
            if (COARSE) {
                double di[GS], dj[GS];
#pragma unroll
                for (int u = 0; u < GS; ++u) {
                    di[u] = fr[u] - cr[u];
                    dj[u] = fcol[u] - ccol[u];
                }
                MB_FOR_BLOCKS(dmma(av[u][v][0], av[u][v][1], di[u] * di[u], fcol[v] * fcol[v]);)
                MB_FOR_BLOCKS(dmma(av[u][v][0], av[u][v][1], cr[u] * cr[u], dj[v] * dj[v]);)
                MB_FOR_BLOCKS(dmma(av[u][v][0], av[u][v][1], 2.0 * (di[u] * cr[u]), fcol[v] * dj[v]);)
            } else {
                MB_FOR_BLOCKS(dmma(av[u][v][0], av[u][v][1], fr[u] * fr[u], fcol[v] * fcol[v]);)
            }
        }
    }
#undef MB_FOR_BLOCKS
}

// MODE 0: covariance sums only; 1: covariance sums + sums of squares; 2: Gram of the differences
template <bool COARSE, int MODE, int GS>
__global__ void __launch_bounds__(kThreadsGram, 1) gram_kernel(const GramArgs a) {
    extern __shared__ double sm[];
    const GramPlan& pl = a.plan;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int LD = pl.ld, NS = pl.ns, nb = pl.nb, r_pad = 8 * nb;
    double* const phi_f = sm;
    double* const phi_c = sm + (size_t)NS * LD;
    int* const flags = reinterpret_cast<int*>(sm + (size_t)2 * NS * LD);   // [NS] sample validity
    __shared__ unsigned cnt_sm[2];
    if (tid == 0) cnt_sm[0] = cnt_sm[1] = 0;

    const int my_tasks = pl.n_tasks[warp];
    double acc_m[kMaxSlots][GS][GS][2];
    double acc_v[MODE == 1 ? kMaxSlots : 1][GS][GS][2];
#pragma unroll
    for (int s = 0; s < kMaxSlots; ++s)
#pragma unroll
        for (int u = 0; u < GS; ++u)
#pragma unroll
            for (int v = 0; v < GS; ++v) {
                acc_m[s][u][v][0] = acc_m[s][u][v][1] = 0.0;
                if (MODE == 1) acc_v[s][u][v][0] = acc_v[s][u][v][1] = 0.0;
            }

    const int n_sides = COARSE ? 2 : 1;
    const int64_t n_tiles = (a.n + NS - 1) / NS;
    const int frag_off = (lane & 3) * LD + (lane >> 2);
    unsigned cnt_ok = 0, cnt_rm = 0;

    // decode this warp's task list ONCE: shared-memory column offsets of its row / column fragments and the SHAPE of
    // each task = which of its GS x GS 8x8 blocks exist and lie on or above the diagonal:
    //   0 = all, 1 = first column only (last, half-filled column group), 2 = upper triangle (diagonal task),
    //   3 = block (0,0) only, 4 = empty slot.
    int off_r[kMaxSlots][GS], off_c[kMaxSlots][GS], shape[kMaxSlots];
#pragma unroll
    for (int slot = 0; slot < kMaxSlots; ++slot) {
        const bool used = slot < my_tasks;
        const int gi = used ? pl.gi[warp][slot] : 0, gj = used ? pl.gj[warp][slot] : 0;
        const int bi0 = gi * GS, bj0 = gj * GS;
#pragma unroll
        for (int u = 0; u < GS; ++u) {
            off_r[slot][u] = 8 * min(bi0 + u, nb - 1);
            off_c[slot][u] = 8 * min(bj0 + u, nb - 1);
        }
        const bool col_full = bj0 + GS <= nb;                   // the row group of a task is full unless gi == gj
        shape[slot] = !used ? 4 : (gi == gj ? ((GS > 1 && col_full) ? 2 : 3) : ((GS > 1 && col_full) ? 0 : 1));
    }

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t n0 = tile * NS;
        // ---- phase A: basis rows of the tile into shared memory ----
        for (int s = tid; s < NS; s += kThreadsGram) flags[s] = 1;
        __syncthreads();
        // pass 1: map values, AND the validity of both sides
        for (int w = tid; w < NS * n_sides; w += kThreadsGram) {
            const int side = w >= NS ? 1 : 0, s = w - side * NS;
            const int64_t n = n0 + s;
            if (n < a.n) {
                const double x = __ldcs(a.pairs + n * a.stride_n + side * a.stride_side);
                const double t = a.basis.kind == MLMCB200_RAW ? x : map_to_ref(a.basis, x);
                if (!moments_finite(a.basis, t)) flags[s] = 0;
                (side == 0 ? phi_f : phi_c)[(size_t)s * LD] = t;           // park t in column 0
            } else {
                flags[s] = 0;
            }
        }
        __syncthreads();
        for (int w = tid; w < NS * n_sides; w += kThreadsGram) {
            const int side = w >= NS ? 1 : 0, s = w - side * NS;
            double* row = (side == 0 ? phi_f : phi_c) + (size_t)s * LD;
            const bool good = flags[s] != 0;
            if (side == 0 && n0 + s < a.n) {
                cnt_ok += good ? 1u : 0u;
                cnt_rm += good ? 0u : 1u;
            }
            write_row(a.basis, row[0], good, row, r_pad);
        }
        __syncthreads();

        // ---- phase B: DMMA over the tile, 4 samples per step ----
        for (int k0 = 0; k0 < NS; k0 += 4) {
            const double* pf = phi_f + (size_t)k0 * LD + frag_off;
            const double* pc = phi_c + (size_t)k0 * LD + frag_off;
#pragma unroll
            for (int slot = 0; slot < kMaxSlots; ++slot) {
                const double* rf = pf;
                const double* rc = pc;
                switch (shape[slot]) {                         // warp-uniform; straight-line DMMA runs inside
                    case 0: slot_mma<COARSE, MODE, GS, 0>(acc_m[slot], acc_v[MODE == 1 ? slot : 0], rf, rc, off_r[slot], off_c[slot]); break;
                    case 1: slot_mma<COARSE, MODE, GS, 1>(acc_m[slot], acc_v[MODE == 1 ? slot : 0], rf, rc, off_r[slot], off_c[slot]); break;
                    case 2: slot_mma<COARSE, MODE, GS, 2>(acc_m[slot], acc_v[MODE == 1 ? slot : 0], rf, rc, off_r[slot], off_c[slot]); break;
                    case 3: slot_mma<COARSE, MODE, GS, 3>(acc_m[slot], acc_v[MODE == 1 ? slot : 0], rf, rc, off_r[slot], off_c[slot]); break;
                    default: break;
                }
            }
        }
        __syncthreads();
    }

    // ---- epilogue: one partial [2 + 2 R R] per CTA, upper blocks mirrored ----
    const int R = a.basis.size;
    double* const out = a.partial + (int64_t)blockIdx.x * a.partial_stride;
    double* const out_m = out + 2;
    double* const out_v = out + 2 + (int64_t)R * R;
#pragma unroll
    for (int slot = 0; slot < kMaxSlots; ++slot) {
        if (slot < my_tasks) {
            const int bi0 = pl.gi[warp][slot] * GS, bj0 = pl.gj[warp][slot] * GS;
#pragma unroll
            for (int u = 0; u < GS; ++u)
#pragma unroll
                for (int v = 0; v < GS; ++v) {
                    const int I = bi0 + u, J = bj0 + v;
                    if (I < nb && J < nb && J >= I) {
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const int i = 8 * I + (lane >> 2), j = 8 * J + 2 * (lane & 3) + e;
                            // diagonal blocks: keep the upper triangle and mirror it, so the result is exactly symmetric
                            if (i < R && j < R && (I != J || j >= i)) {
                                out_m[(int64_t)i * R + j] = acc_m[slot][u][v][e];
                                if (i != j) out_m[(int64_t)j * R + i] = acc_m[slot][u][v][e];
                                if (MODE == 1) {
                                    out_v[(int64_t)i * R + j] = acc_v[slot][u][v][e];
                                    if (i != j) out_v[(int64_t)j * R + i] = acc_v[slot][u][v][e];
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

int make_plan(int R, GramPlan* pl, size_t* smem) {
    const int nb = (R + 7) / 8;
    pl->nb = nb;
    pl->gs = nb <= 5 ? 1 : 2;
    const int ng = (nb + pl->gs - 1) / pl->gs;
    const int n_tasks = ng * (ng + 1) / 2;
    if (n_tasks > kWarps * kMaxSlots) {
        set_error("gram: %d moments need %d block tasks (max %d)", R, n_tasks, kWarps * kMaxSlots);
        return -1;
    }
    // task weights = number of 8x8 blocks on/above the diagonal, assigned greedily (largest first)
    struct T { int gi, gj, w; } tasks[kWarps * kMaxSlots];
    int nt = 0;
    for (int gi = 0; gi < ng; ++gi)
        for (int gj = gi; gj < ng; ++gj) {
            int w = 0;
            for (int u = 0; u < pl->gs; ++u)
                for (int v = 0; v < pl->gs; ++v) {
                    const int I = gi * pl->gs + u, J = gj * pl->gs + v;
                    if (I < nb && J < nb && J >= I) ++w;
                }
            tasks[nt++] = {gi, gj, w};
        }
    for (int i = 0; i < nt; ++i)
        for (int j = i + 1; j < nt; ++j)
            if (tasks[j].w > tasks[i].w) { T t = tasks[i]; tasks[i] = tasks[j]; tasks[j] = t; }
    int load[kWarps] = {0};
    for (int w = 0; w < kWarps; ++w) pl->n_tasks[w] = 0;
    for (int i = 0; i < nt; ++i) {
        int best = -1;
        for (int w = 0; w < kWarps; ++w)
            if (pl->n_tasks[w] < kMaxSlots && (best < 0 || load[w] < load[best])) best = w;
        pl->gi[best][pl->n_tasks[best]] = (unsigned char)tasks[i].gi;
        pl->gj[best][pl->n_tasks[best]] = (unsigned char)tasks[i].gj;
        pl->n_tasks[best]++;
        load[best] += tasks[i].w;
    }
    int ld = 8 * nb;
    while (ld % 8 != 4) ++ld;
    pl->ld = ld;
    const size_t budget = 200u * 1024u;
    int ns = (int)(budget / ((size_t)2 * ld * sizeof(double)));
    ns = (ns / 32) * 32;
    if (ns > 128) ns = 128;
    if (ns < 32) {
        set_error("gram: %d moments do not fit the shared-memory tile", R);
        return -1;
    }
    pl->ns = ns;
    *smem = (size_t)2 * ns * ld * sizeof(double) + (size_t)ns * sizeof(int);
    return 0;
}

template <bool COARSE, int MODE, int GS>
int launch_gram(const GramArgs& a, int grid, size_t smem, cudaStream_t st) {
    auto kern = gram_kernel<COARSE, MODE, GS>;
    MB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreadsGram, smem, st>>>(a);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}

template <bool COARSE, int MODE>
int launch_gram_gs(const GramArgs& a, int grid, size_t smem, cudaStream_t st) {
    return a.plan.gs == 1 ? launch_gram<COARSE, MODE, 1>(a, grid, smem, st)
                          : launch_gram<COARSE, MODE, 2>(a, grid, smem, st);
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
    double g_acc = 0.0, f_acc = 0.0;                            // thread tid < R owns g_tid; thread 0 owns F
    const int frag_off = (lane & 3) * LD + (lane >> 2);
    const int64_t n_tiles = (a.n_nodes + NS - 1) / NS;

    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t q0 = tile * NS;
        __syncthreads();
        // stage NS rows (zero-padded columns / rows)
        for (int idx = tid; idx < NS * r_pad; idx += kThreadsGram) {
            const int s = idx / r_pad, i = idx - s * r_pad;
            const int64_t q = q0 + s;
            phi_s[(size_t)s * LD + i] = (q < a.n_nodes && i < R) ? __ldg(a.phi + q * a.ld_g + i) : 0.0;
        }
        __syncthreads();
        for (int s = warp; s < NS; s += kWarps) {
            double dot = 0.0;
            for (int i = lane; i < R; i += 32) dot = fma(phi_s[(size_t)s * LD + i], lam_s[i], dot);
            dot = warp_sum(dot);
            if (lane == 0) {
                const double power = fmin(fmax(-dot, -200.0), 200.0);
                wr[s] = (q0 + s < a.n_nodes) ? a.w[q0 + s] * exp(power) : 0.0;
            }
        }
        __syncthreads();
        if (tid < R) {
            for (int s = 0; s < NS; ++s) g_acc = fma(wr[s], phi_s[(size_t)s * LD + tid], g_acc);
        }
        if (tid == kThreadsGram - 1) {
            for (int s = 0; s < NS; ++s) f_acc += wr[s];
        }
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
                                if (I < nb && J < nb && J >= I) dmma(acc[slot][u][v][0], acc[slot][u][v][1], fr[u], fcol[v]);
                            }
                    }
                }
            }
        }
    }

    double* const out = a.partial + (int64_t)blockIdx.x * a.partial_stride;
    if (tid == kThreadsGram - 1) out[0] = f_acc;
    if (tid < R) out[1 + tid] = g_acc;
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
                        if (I < nb && J < nb && J >= I) {
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

// out[j] = sum_b partial[b][j]
__global__ void sum_partials_kernel(const double* __restrict__ partial, int n_partials, int64_t stride, int64_t len,
                                    double* __restrict__ out) {
    const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= len) return;
    double s = 0.0;
    for (int b = 0; b < n_partials; ++b) s += partial[(int64_t)b * stride + j];
    out[j] = s;
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
    if (size < 1 || size > MLMCB200_MAX_MOMENTS) return -1;
    return (int64_t)sm_count() * (2 + 2 * (int64_t)size * size) * (int64_t)sizeof(double);
}

extern "C" int mlmcb200_gram_accumulate(const mlmcb200_basis_t* basis, const double* pairs, int64_t n,
                                        int64_t stride_n, int64_t stride_side, int32_t has_coarse,
                                        int32_t mode, int32_t want_var, double* acc,
                                        void* workspace, int64_t workspace_bytes, void* stream) {
    if (check_basis(basis) != 0) return -1;
    MB_REQUIRE(n >= 0 && (mode == 0 || mode == 1), "gram_accumulate: bad n=%lld mode=%d", (long long)n, mode);
    MB_REQUIRE(acc != nullptr && workspace != nullptr, "gram_accumulate: null acc/workspace");
    MB_REQUIRE(mode == 0 || has_coarse, "gram_accumulate: mode 1 (difference Gram) needs a coarse side");
    if (n == 0) return 0;
    MB_REQUIRE(pairs != nullptr, "gram_accumulate: null pairs");
    GramArgs a;
    a.basis = *basis;
    a.pairs = pairs;
    a.n = n;
    a.stride_n = stride_n;
    a.stride_side = stride_side;
    size_t smem = 0;
    if (make_plan(basis->size, &a.plan, &smem) != 0) return -1;
    const int64_t R2 = (int64_t)basis->size * basis->size;
    const int64_t stride = 2 + 2 * R2;
    const int64_t tiles = (n + a.plan.ns - 1) / a.plan.ns;
    int grid = sm_count();
    if (tiles < grid) grid = (int)tiles;
    MB_REQUIRE(workspace_bytes >= (int64_t)grid * stride * 8, "gram_accumulate: workspace too small");
    a.partial = static_cast<double*>(workspace);
    a.partial_stride = stride;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (mode == 1)
        rc = launch_gram_gs<true, 2>(a, grid, smem, st);
    else if (want_var)
        rc = has_coarse ? launch_gram_gs<true, 1>(a, grid, smem, st) : launch_gram_gs<false, 1>(a, grid, smem, st);
    else
        rc = has_coarse ? launch_gram_gs<true, 0>(a, grid, smem, st) : launch_gram_gs<false, 0>(a, grid, smem, st);
    if (rc != 0) return rc;
    // sums always; sums of squares only when they were produced
    const int64_t len = want_var && mode == 0 ? stride : 2 + R2;
    return launch_reduce_partials(a.partial, grid, stride, len, acc, st);
}

extern "C" int64_t mlmcb200_maxent_workspace_bytes(int64_t n_nodes, int32_t size) {
    if (size < 1 || size > MLMCB200_MAX_MOMENTS || n_nodes < 0) return -1;
    return (int64_t)sm_count() * (1 + (int64_t)size + (int64_t)size * size) * (int64_t)sizeof(double);
}

extern "C" int mlmcb200_maxent_fgh(const double* phi, int64_t ld, const double* w, const double* lam_scaled,
                                   int64_t n_nodes, int32_t size, int32_t what, double* out,
                                   void* workspace, int64_t workspace_bytes, void* stream) {
    MB_REQUIRE(phi && w && lam_scaled && out && workspace, "maxent_fgh: null pointer");
    MB_REQUIRE(size >= 1 && size <= MLMCB200_MAX_MOMENTS && ld >= size && n_nodes >= 1, "maxent_fgh: bad sizes");
    MB_REQUIRE((what & 7) != 0, "maxent_fgh: nothing requested");
    MaxentArgs a;
    a.phi = phi;
    a.ld_g = ld;
    a.w = w;
    a.lam = lam_scaled;
    a.n_nodes = n_nodes;
    a.R = size;
    a.want_h = (what & 4) ? 1 : 0;
    size_t smem = 0;
    if (make_plan(size, &a.plan, &smem) != 0) return -1;
    // single table + weights + multipliers instead of two tables + flags
    const int ns = a.plan.ns;
    smem = ((size_t)ns * a.plan.ld + ns + 8 * a.plan.nb) * sizeof(double);
    const int grid = maxent_grid(n_nodes, ns);
    const int64_t stride = 1 + (int64_t)size + (int64_t)size * size;
    MB_REQUIRE(workspace_bytes >= (int64_t)grid * stride * 8, "maxent_fgh: workspace too small");
    a.partial = static_cast<double*>(workspace);
    a.partial_stride = stride;
    cudaStream_t st = (cudaStream_t)stream;
    auto kern = a.plan.gs == 1 ? maxent_kernel<1> : maxent_kernel<2>;
    MB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kThreadsGram, smem, st>>>(a);
    MB_CUDA_OK(cudaGetLastError());
    const int64_t len = a.want_h ? stride : 1 + size;
    const int threads = 256;
    sum_partials_kernel<<<(unsigned)((len + threads - 1) / threads), threads, 0, st>>>(a.partial, grid, stride, len, out);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}
