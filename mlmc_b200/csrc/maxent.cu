// Max-entropy functional pieces on a fixed node set, up to 64 moments (mlmc/tool/simple_distribution.py:254-327):
//     rho_q = exp(clip(-Phi_q . lam, -200, 200)),  F = sum_q w_q rho_q,  g = Phi^T (w rho),  H = Phi^T diag(w rho) Phi
//
// The first version (maxent_kernel in gram.cu, still used above 64 moments) did everything in one CTA-per-SM kernel whose
// tile staging was a serial chain per node (global load -> shuffle reduction -> exp): 0.11 ms per evaluation at cfg4's
// size (100 002 nodes x 50 moments) against ~10 us of FP64 tensor-pipe time.  Now an evaluation is two throughput-bound
// kernels and one small reduction:
//   1. maxent_rho_kernel : 4 lanes per node, the node's row stays in REGISTERS: dot product with the multipliers (two
//      quad shuffles), exponent, weight w rho -> global, and the F / g partial sums from the same registers.  No shared
//      memory in the loop, ~2.6 nodes per quad at cfg4's size: one pass at full occupancy.
//   2. maxent_h_kernel   : H = (diag(w rho) Phi)^T Phi on DMMA m8n8k4.  Every CTA owns an equal, contiguous share of the
//      nodes; a 64-node tile (Phi and its weighted copy) is written by all threads from registers that were fetched one
//      tile ahead (column-major thread mapping: no index arithmetic, coalesced); the upper-triangle 8x8 blocks are dealt
//      to 4 groups of 4 warps and the warps of a group split the samples of the tile (split-K): every warp issues the
//      same, straight-line sequence of DMMAs (number of blocks per side = template parameter: all tile offsets are
//      immediates); the partial accumulators of a block meet in shared memory at the end, in a fixed order.  Two CTAs of 8 warps per SM:
//      one contracts while the other waits for its next tile.
//   3. maxent_sum_kernel : fixed-order sums of the per-CTA partials -> out = [F | g | H].
#include "common.cuh"

namespace mlmcb200 {
namespace {

constexpr int kMeWarps = 8;
constexpr int kMeThreads = kMeWarps * 32;
constexpr int kMeGroups = 4;            // block groups; kMeWarps / kMeGroups warps split the samples of a tile
constexpr int kMeSplit = kMeWarps / kMeGroups;
constexpr int kMeCtasPerSm = 2;         // two CTAs per SM: one contracts while the other waits for its next tile
constexpr int kRhoThreads = 256;

// ---- 1. exponents, weights, F and g ----
// NI: row elements per lane (lane p of a quad holds columns p, p + 4, ...)
template <int NI>
__global__ void __launch_bounds__(kRhoThreads) maxent_rho_kernel(const double* __restrict__ phi, int64_t ld_g,
                                                                 const double* __restrict__ w,
                                                                 const double* __restrict__ lam, int64_t n_nodes, int R,
                                                                 double* __restrict__ wr_out,
                                                                 double* __restrict__ partial_fg) {
    __shared__ double red[kRhoThreads / 32][4 * NI + 1];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, p = lane & 3;
    double lam_r[NI], g_loc[NI], f_loc = 0.0;
#pragma unroll
    for (int j = 0; j < NI; ++j) {
        const int i = p + 4 * j;
        lam_r[j] = i < R ? __ldg(lam + i) : 0.0;
        g_loc[j] = 0.0;
    }
    const int64_t n_slots = (int64_t)gridDim.x * (kRhoThreads / 4);
    // whole warps iterate together (the quad shuffles need all lanes): the bound is rounded up to 8 nodes per warp
    for (int64_t q0 = ((int64_t)blockIdx.x * kRhoThreads + (tid & ~31)) / 4; q0 < n_nodes; q0 += n_slots) {
        const int64_t q = q0 + (lane >> 2);
        const bool in = q < n_nodes;
        double v[NI];
        double dot = 0.0;
#pragma unroll
        for (int j = 0; j < NI; ++j) {
            const int i = p + 4 * j;
            v[j] = (in && i < R) ? __ldg(phi + q * ld_g + i) : 0.0;
        }
        const double wq0 = in ? __ldg(w + q) : 0.0;
#pragma unroll
        for (int j = 0; j < NI; ++j) dot = fma(v[j], lam_r[j], dot);
        dot += __shfl_xor_sync(0xffffffffu, dot, 1);
        dot += __shfl_xor_sync(0xffffffffu, dot, 2);
        const double wq = wq0 * exp(fmin(fmax(-dot, -200.0), 200.0));
        if (p == 0 && in) {
            wr_out[q] = wq;
            f_loc += wq;
        }
#pragma unroll
        for (int j = 0; j < NI; ++j) g_loc[j] = fma(wq, v[j], g_loc[j]);
    }
    // lanes with the same p hold the same columns: fold the 8 quads of the warp, then the warps of the CTA
#pragma unroll
    for (int j = 0; j < NI; ++j) {
        double s = g_loc[j];
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        s += __shfl_xor_sync(0xffffffffu, s, 8);
        s += __shfl_xor_sync(0xffffffffu, s, 16);
        if (lane < 4) red[warp][p + 4 * j] = s;
    }
    f_loc += __shfl_xor_sync(0xffffffffu, f_loc, 4);
    f_loc += __shfl_xor_sync(0xffffffffu, f_loc, 8);
    f_loc += __shfl_xor_sync(0xffffffffu, f_loc, 16);
    if (lane == 0) red[warp][4 * NI] = f_loc;
    __syncthreads();
    double* const out = partial_fg + (int64_t)blockIdx.x * (1 + R);
    for (int i = tid; i < 1 + R; i += kRhoThreads) {
        double s = 0.0;
        const int col = i == 0 ? 4 * NI : i - 1;
        for (int wv = 0; wv < kRhoThreads / 32; ++wv) s += red[wv][col];
        out[i] = s;
    }
}

// ---- 2. H ----
constexpr int kMeNS = 64;               // nodes per tile
constexpr int kMeMaxBlocks = 12;        // 8x8 blocks per warp group (array bound)

struct MaxentHArgs {
    const double* phi;
    int64_t ld_g;
    const double* wr;       // w rho of every node (kernel 1)
    int64_t n_nodes;
    int64_t nodes_per_cta;
    int R;
    double* partial;        // [gridDim.x][R R]
    int n_blocks[kMeGroups];
    unsigned new_row[kMeGroups];   // bit b: block b of the group starts a new block row (its A fragment must be loaded)
    unsigned char bi[kMeGroups][kMeMaxBlocks];
    unsigned char bj[kMeGroups][kMeMaxBlocks];
};

__device__ __forceinline__ void dmma_me(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// NB: 8x8 blocks per side (compile time: the shared-memory row stride LD = 8 NB + 4 and every fragment offset inside
// the tile loop are immediates).  Thread (ci = tid % 64, s0 = tid / 64) owns column ci of the rows s0, s0 + 8, ... of
// every tile.  A warp's block list is padded to NBW entries with copies of block (0, 0) whose accumulators are never
// written out: the DMMA sequence is straight-line, no predication around mma.sync.
template <int NB>
__global__ void __launch_bounds__(kMeThreads, kMeCtasPerSm) maxent_h_kernel(const MaxentHArgs a) {
    constexpr int LD = 8 * NB + 4, NS = kMeNS, r_pad = 8 * NB;
    constexpr int NBW = (NB * (NB + 1) / 2 + kMeGroups - 1) / kMeGroups;
    constexpr int kRowGroups = kMeThreads / 64;                // thread (ci = tid % 64, s0 = tid / 64): rows s0, s0 + 4, ...
    constexpr int kRows = NS / kRowGroups;                     // rows per thread and tile
    constexpr int kPer = NS / kMeSplit;                        // samples of a warp's share of the tile
    extern __shared__ double sm[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int R = a.R;
    double* const tile = sm;                                   // Phi rows        [NS][LD]
    double* const tile_a = sm + NS * LD;                       // w rho Phi rows  [NS][LD]

    const int64_t q_begin = (int64_t)blockIdx.x * a.nodes_per_cta;
    const int64_t q_end = min(q_begin + a.nodes_per_cta, a.n_nodes);
    const int n_tiles = q_end > q_begin ? (int)((q_end - q_begin + NS - 1) / NS) : 0;

    const int group = warp / kMeSplit, kq = warp % kMeSplit;  // block group, share of the tile's samples
    const int n_blk = a.n_blocks[group];
    const unsigned nr = a.new_row[group];
    const int frag = (kq * kPer + (lane & 3)) * LD + (lane >> 2);
    int oa[NBW], ob[NBW];                                     // fragment offsets inside the tiles
#pragma unroll
    for (int b = 0; b < NBW; ++b) {
        oa[b] = frag + (b < n_blk ? 8 * a.bi[group][b] : 0);
        ob[b] = frag + (b < n_blk ? 8 * a.bj[group][b] : 0);
    }
    double acc[NBW][2];
#pragma unroll
    for (int b = 0; b < NBW; ++b) acc[b][0] = acc[b][1] = 0.0;

    const int ci = tid & 63, s0 = tid >> 6;
    double pre[kRows];                                         // this thread's elements of the NEXT tile
    auto fetch = [&](int t) {
        const int64_t q0 = q_begin + (int64_t)t * NS + s0;
#pragma unroll
        for (int m = 0; m < kRows; ++m) {
            const int64_t q = q0 + kRowGroups * m;
            pre[m] = (t < n_tiles && q < q_end && ci < R) ? __ldg(a.phi + q * a.ld_g + ci) : 0.0;
        }
    };
    fetch(0);
    for (int t = 0; t < n_tiles; ++t) {
        // the weights of the tile's rows come from kernel 1 (L2-resident), 16 independent loads
        double wq[kRows];
        const int64_t q0 = q_begin + (int64_t)t * NS + s0;
#pragma unroll
        for (int m = 0; m < kRows; ++m) {
            const int64_t q = q0 + kRowGroups * m;
            wq[m] = q < q_end ? __ldg(a.wr + q) : 0.0;
        }
        __syncthreads();                                       // the previous tile has been consumed
        if (ci < r_pad) {
#pragma unroll
            for (int m = 0; m < kRows; ++m) {
                tile[(kRowGroups * m + s0) * LD + ci] = pre[m];
                tile_a[(kRowGroups * m + s0) * LD + ci] = wq[m] * pre[m];
            }
        }
        fetch(t + 1);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kPer / 4; ++k) {
            // blocks are listed row by row: consecutive blocks of a row share the A fragment
            double fa[NBW], fb[NBW];
#pragma unroll
            for (int b = 0; b < NBW; ++b) {
                fa[b] = b == 0 ? 0.0 : fa[b - 1];
                if ((nr >> b) & 1u) fa[b] = tile_a[oa[b] + k * 4 * LD];
                fb[b] = tile[ob[b] + k * 4 * LD];
            }
#pragma unroll
            for (int b = 0; b < NBW; ++b) dmma_me(acc[b][0], acc[b][1], fa[b], fb[b]);
        }
    }

    // ---------------- CTA epilogue: fold the four sample quarters of every block, fixed order ----------------
    double* const out_h = a.partial + (int64_t)blockIdx.x * R * R;
    double* const exch = sm;                                   // [kMeWarps][64]
#pragma unroll
    for (int b = 0; b < NBW; ++b) {
        __syncthreads();
        exch[warp * 64 + 2 * lane] = acc[b][0];
        exch[warp * 64 + 2 * lane + 1] = acc[b][1];
        __syncthreads();
        if (kq == 0 && b < n_blk) {
            const int I = a.bi[group][b], J = a.bj[group][b];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                double v = exch[warp * 64 + 2 * lane + e];
                for (int k = 1; k < kMeSplit; ++k) v += exch[(warp + k) * 64 + 2 * lane + e];
                const int i = 8 * I + (lane >> 2), j = 8 * J + 2 * (lane & 3) + e;
                if (i < R && j < R && (I != J || j >= i)) {       // diagonal blocks: upper triangle, mirrored -> symmetric
                    out_h[(int64_t)i * R + j] = v;
                    if (i != j) out_h[(int64_t)j * R + i] = v;
                }
            }
        }
    }
}

template <int NB>
int launch_h(const MaxentHArgs& a, int grid, cudaStream_t st) {
    constexpr size_t smem = (size_t)2 * kMeNS * (8 * NB + 4) * sizeof(double);       // Phi tile + weighted copy
    static_assert(smem >= (size_t)kMeWarps * 64 * sizeof(double), "exchange array fits the tiles");
    auto kern = maxent_h_kernel<NB>;
    MB_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, kMeThreads, smem, st>>>(a);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}

// ---- 3. out[0 .. R] = sum of the F / g partials, out[1 + R ..] = sum of the H partials (fixed order) ----
__global__ void maxent_sum_kernel(const double* __restrict__ partial_fg, int n_fg, const double* __restrict__ partial_h,
                                  int n_h, int R, int want_h, double* __restrict__ out) {
    // one WARP per output: lanes stride over the partials (independent loads), then a shuffle tree -- a thread per output
    // walking ~300 partials one after the other took longer than the two kernels before it
    const int64_t j = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t n_fg_out = 1 + R, total = n_fg_out + (want_h ? (int64_t)R * R : 0);
    if (j >= total) return;
    double s = 0.0;
    if (j < n_fg_out) {
        for (int b = lane; b < n_fg; b += 32) s += partial_fg[(int64_t)b * n_fg_out + j];
    } else {
        const int64_t k = j - n_fg_out;
        for (int b = lane; b < n_h; b += 32) s += partial_h[(int64_t)b * R * R + k];
    }
    s = warp_sum(s);
    if (lane == 0) out[j] = s;
}

int rho_grid(int64_t n_nodes) {
    int64_t blocks = (n_nodes * 4 + kRhoThreads - 1) / kRhoThreads;
    const int64_t cap = (int64_t)sm_count() * 4;
    if (blocks > cap) blocks = cap;
    return (int)(blocks < 1 ? 1 : blocks);
}

}  // namespace

int64_t maxent_fast_workspace_bytes(int64_t n_nodes, int R) {
    const int64_t fg = (int64_t)sm_count() * 4 * (1 + R), h = (int64_t)sm_count() * kMeCtasPerSm * R * R;
    return (n_nodes + fg + h + 8) * (int64_t)sizeof(double);
}

// Returns 1 if these kernels do not cover the size (the caller falls back), 0 on success, < 0 on error.
int maxent_fast_launch(const double* phi, int64_t ld_g, const double* w, const double* lam, int64_t n_nodes, int R,
                       int want_h, double* out, void* workspace, int64_t workspace_bytes, cudaStream_t st) {
    const int nb = (R + 7) / 8;
    const int total_blocks = nb * (nb + 1) / 2;
    if (nb > 8) return 1;                                      // more than 64 moments: the first version (gram.cu)
    MB_REQUIRE(workspace_bytes >= maxent_fast_workspace_bytes(n_nodes, R), "maxent_fgh: workspace too small");
    double* const wr = static_cast<double*>(workspace);
    double* const partial_fg = wr + n_nodes;
    double* const partial_h = partial_fg + (int64_t)sm_count() * 4 * (1 + R);

    const int grid_rho = rho_grid(n_nodes);
    if (nb <= 4)
        maxent_rho_kernel<8><<<grid_rho, kRhoThreads, 0, st>>>(phi, ld_g, w, lam, n_nodes, R, wr, partial_fg);
    else
        maxent_rho_kernel<16><<<grid_rho, kRhoThreads, 0, st>>>(phi, ld_g, w, lam, n_nodes, R, wr, partial_fg);
    MB_CUDA_OK(cudaGetLastError());

    int grid_h = 0;
    if (want_h) {
        MaxentHArgs a;
        a.phi = phi;
        a.ld_g = ld_g;
        a.wr = wr;
        a.n_nodes = n_nodes;
        a.R = R;
        a.partial = partial_h;
        // upper-triangle blocks row by row, dealt to the four warp groups in runs
        const int per = (total_blocks + kMeGroups - 1) / kMeGroups;
        for (int gI = 0; gI < kMeGroups; ++gI) a.n_blocks[gI] = 0, a.new_row[gI] = 0;
        int idx = 0;
        for (int I = 0; I < nb; ++I)
            for (int J = I; J < nb; ++J, ++idx) {
                const int gI = idx / per, k = a.n_blocks[gI];
                a.bi[gI][k] = (unsigned char)I;
                a.bj[gI][k] = (unsigned char)J;
                if (k == 0 || a.bi[gI][k - 1] != I) a.new_row[gI] |= 1u << k;
                ++a.n_blocks[gI];
            }
        for (int gI = 0; gI < kMeGroups; ++gI) a.new_row[gI] |= 1u;      // empty groups still load a (dummy) fragment
        grid_h = sm_count() * kMeCtasPerSm;
        if ((int64_t)grid_h * kMeNS > n_nodes) grid_h = (int)((n_nodes + kMeNS - 1) / kMeNS);   // >= one tile per CTA
        if (grid_h < 1) grid_h = 1;
        a.nodes_per_cta = (n_nodes + grid_h - 1) / grid_h;
        int rc = -1;
        switch (nb) {
            case 1: rc = launch_h<1>(a, grid_h, st); break;
            case 2: rc = launch_h<2>(a, grid_h, st); break;
            case 3: rc = launch_h<3>(a, grid_h, st); break;
            case 4: rc = launch_h<4>(a, grid_h, st); break;
            case 5: rc = launch_h<5>(a, grid_h, st); break;
            case 6: rc = launch_h<6>(a, grid_h, st); break;
            case 7: rc = launch_h<7>(a, grid_h, st); break;
            case 8: rc = launch_h<8>(a, grid_h, st); break;
        }
        if (rc != 0) return rc;
    }
    const int64_t total = 1 + R + (want_h ? (int64_t)R * R : 0);
    maxent_sum_kernel<<<(unsigned)((total * 32 + 255) / 256), 256, 0, st>>>(partial_fg, grid_rho, partial_h, grid_h, R,
                                                                             want_h, out);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}

}  // namespace mlmcb200
