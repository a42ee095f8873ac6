// Moments.eval_all on the device: x[n] -> Phi[n][R] (mlmc/moments.py:75-93; Monomial :122-126, Fourier :145-162,
// Legendre :195-197, TransformedMoments :256-259).
//
// Not the throughput path (the fused kernels never materialise Phi); it serves Moments.__call__/eval_all,
// SimpleDistribution's quadrature table and density().  Arithmetic follows numpy's legvander / polyvander
// operation by operation (no FMA contraction), so Legendre and Monomial tables are bit-identical to the
// reference's.  One warp evaluates 32 values into a shared-memory tile (odd leading dimension, conflict-free),
// then the tile is written out with lane-contiguous addresses.
#include "common.cuh"

namespace mlmcb200 {
namespace {

struct EvalArgs {
    mlmcb200_basis_t basis;
    const double* x;
    int64_t n;
    const double* matrix;   // [n_rows][size] or null
    int n_rows;
    int n_fn;               // number of base functions to evaluate
    int n_out;              // row length of out
    double* out;
    int ld;                 // tile leading dimension (odd)
};

__global__ void basis_eval_kernel(const EvalArgs a) {
    extern __shared__ double sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    const int R = a.n_fn;
    double* tile = sm + (size_t)warp * 32 * a.ld;
    double* lt = sm + (size_t)n_warps * 32 * a.ld;          // matrix^T : [R][n_out]
    if (a.matrix != nullptr) {
        for (int idx = threadIdx.x; idx < R * a.n_out; idx += blockDim.x) {
            const int i = idx / a.n_out, k = idx - i * a.n_out;
            lt[idx] = a.matrix[(int64_t)k * a.basis.size + i];
        }
        __syncthreads();
    }
    const int64_t n_tiles = (a.n + 31) / 32;
    for (int64_t tile_id = (int64_t)blockIdx.x * n_warps + warp; tile_id < n_tiles;
         tile_id += (int64_t)gridDim.x * n_warps) {
        const int64_t n0 = tile_id * 32;
        const int64_t n = n0 + lane;
        double* row = tile + lane * a.ld;
        if (n < a.n) {
            const double xv = a.x[n];
            if (a.basis.kind == MLMCB200_RAW) {
                row[0] = xv;
            } else {
                const double t = map_to_ref(a.basis, xv);
                if (a.basis.kind == MLMCB200_FOURIER) {
                    row[0] = 1.0;
                    for (int i = 1; i < R; ++i) {
                        const int k = (i + 1) >> 1;
                        const double kt = __dmul_rn(t, (double)k);
                        row[i] = (i & 1) ? cos(kt) : sin(kt);
                    }
                } else {
                    double p0 = __dadd_rn(__dmul_rn(t, 0.0), 1.0);      // numpy: x*0 + 1
                    row[0] = p0;
                    if (R > 1) {
                        double p1 = t;
                        row[1] = p1;
                        for (int i = 2; i < R; ++i) {
                            const double p2 = a.basis.kind == MLMCB200_LEGENDRE ? legendre_step_exact(p1, p0, t, i)
                                                                                : __dmul_rn(p1, t);
                            row[i] = p2;
                            p0 = p1;
                            p1 = p2;
                        }
                    }
                }
            }
        }
        __syncwarp();
        const int rows_here = (int)min((int64_t)32, a.n - n0);
        double* dst = a.out + n0 * a.n_out;
        if (a.matrix == nullptr) {
            const int total = rows_here * R;
            for (int f = lane; f < total; f += 32) {
                const int s = f / R, i = f - s * R;
                dst[f] = tile[s * a.ld + i];
            }
        } else {
            for (int s = 0; s < rows_here; ++s) {
                for (int k = lane; k < a.n_out; k += 32) {
                    double acc = 0.0;
                    for (int i = 0; i < R; ++i) acc = fma(tile[s * a.ld + i], lt[i * a.n_out + k], acc);
                    dst[(int64_t)s * a.n_out + k] = acc;
                }
            }
        }
        __syncwarp();
    }
}

// SimpleDistribution.density (mlmc/tool/simple_distribution.py:96-105): exp(clip(-phi(x) . coef, -200, 200)) without the
// table: one thread per value walks the recurrence (the table's own operations) and accumulates the dot product.
__global__ void density_eval_kernel(const mlmcb200_basis_t basis, const double* __restrict__ x, int64_t n,
                                    const double* __restrict__ coef, int n_coef, double* __restrict__ out) {
    extern __shared__ double cf[];
    for (int i = threadIdx.x; i < n_coef; i += blockDim.x) cf[i] = coef[i];
    __syncthreads();
    for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
        const double xv = x[k];
        double dot;
        if (basis.kind == MLMCB200_RAW) {
            dot = cf[0] * xv;
        } else {
            const double t = map_to_ref(basis, xv);
            if (basis.kind == MLMCB200_FOURIER) {
                dot = cf[0];
                for (int i = 1; i < n_coef; ++i) {
                    const double kt = __dmul_rn(t, (double)((i + 1) >> 1));
                    dot = fma(cf[i], (i & 1) ? cos(kt) : sin(kt), dot);
                }
            } else {
                double p0 = __dadd_rn(__dmul_rn(t, 0.0), 1.0), p1 = t;
                dot = cf[0] * p0;
                if (n_coef > 1) dot = fma(cf[1], p1, dot);
                for (int i = 2; i < n_coef; ++i) {
                    const double p2 = basis.kind == MLMCB200_LEGENDRE ? legendre_step_exact(p1, p0, t, i)
                                                                       : __dmul_rn(p1, t);
                    dot = fma(cf[i], p2, dot);
                    p0 = p1;
                    p1 = p2;
                }
            }
        }
        // np.minimum(np.maximum(power, -200), 200) keeps NaN (a value outside a clipped domain); fmin / fmax would not
        const double power = isnan(dot) ? dot : fmin(fmax(-dot, -200.0), 200.0);
        out[k] = exp(power);
    }
}

}  // namespace
}  // namespace mlmcb200

using namespace mlmcb200;

extern "C" int mlmcb200_density_eval(const mlmcb200_basis_t* basis, const double* x, int64_t n, const double* coef,
                                     int32_t n_coef, double* out, void* stream) {
    if (check_basis(basis) != 0) return -1;
    MB_REQUIRE(n >= 0 && n_coef >= 1 && n_coef <= basis->size, "density_eval: bad n=%lld n_coef=%d (basis size %d)",
               (long long)n, n_coef, basis->size);
    if (n == 0) return 0;
    MB_REQUIRE(x != nullptr && coef != nullptr && out != nullptr, "density_eval: null pointer");
    const int threads = 128;
    int64_t blocks = (n + threads - 1) / threads;
    const int64_t max_blocks = (int64_t)sm_count() * 16;
    if (blocks > max_blocks) blocks = max_blocks;
    density_eval_kernel<<<(unsigned)blocks, threads, (size_t)n_coef * sizeof(double), (cudaStream_t)stream>>>(
        *basis, x, n, coef, n_coef, out);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}

extern "C" int mlmcb200_basis_eval(const mlmcb200_basis_t* basis, const double* x, int64_t n,
                                   const double* matrix, int32_t n_rows, int32_t n_out,
                                   double* out, void* stream) {
    if (check_basis(basis) != 0) return -1;
    MB_REQUIRE(n >= 0 && n_out >= 1, "basis_eval: bad n=%lld n_out=%d", (long long)n, n_out);
    if (n == 0) return 0;
    MB_REQUIRE(x != nullptr && out != nullptr, "basis_eval: null pointer");
    EvalArgs a;
    a.basis = *basis;
    a.x = x;
    a.n = n;
    a.matrix = matrix;
    a.n_rows = n_rows;
    a.out = out;
    a.n_out = n_out;
    if (matrix != nullptr) {
        MB_REQUIRE(n_rows >= n_out, "basis_eval: n_out=%d exceeds matrix rows=%d", n_out, n_rows);
        a.n_fn = basis->size;
    } else {
        MB_REQUIRE(n_out <= basis->size, "basis_eval: n_out=%d exceeds basis size=%d", n_out, basis->size);
        a.n_fn = n_out;
    }
    a.ld = a.n_fn | 1;
    int n_warps = 4;
    size_t lt_bytes = matrix ? (size_t)a.n_fn * n_out * 8 : 0;
    size_t smem = 0;
    for (; n_warps >= 1; n_warps >>= 1) {
        smem = (size_t)n_warps * 32 * a.ld * 8 + lt_bytes;
        if (smem <= 227u * 1024u) break;
    }
    MB_REQUIRE(n_warps >= 1, "basis_eval: %d x %d does not fit shared memory", a.n_fn, n_out);
    const int64_t tiles = (n + 31) / 32;
    int64_t blocks = (tiles + n_warps - 1) / n_warps;
    const int64_t max_blocks = (int64_t)sm_count() * 8;
    if (blocks > max_blocks) blocks = max_blocks;
    MB_CUDA_OK(cudaFuncSetAttribute(basis_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    basis_eval_kernel<<<(unsigned)blocks, n_warps * 32, smem, (cudaStream_t)stream>>>(a);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}
