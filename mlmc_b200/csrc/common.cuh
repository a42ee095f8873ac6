// Shared host/device helpers of libmlmcb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/mlmcb200.h"

namespace mlmcb200 {

void set_error(const char* fmt, ...);
int  sm_count();

#define MB_CUDA_OK(expr)                                                                     \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::mlmcb200::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),    \
                                  __FILE__, __LINE__);                                       \
            return -2;                                                                       \
        }                                                                                    \
    } while (0)

#define MB_REQUIRE(cond, ...)                                                                \
    do {                                                                                     \
        if (!(cond)) {                                                                       \
            ::mlmcb200::set_error(__VA_ARGS__);                                              \
            return -1;                                                                       \
        }                                                                                    \
    } while (0)

int check_basis(const mlmcb200_basis_t* b);

// Coefficients of the monic Legendre recurrence W_i = t W_{i-1} - kLegCoef[i] W_{i-2}, P_i = kLegAlpha[i] W_i
// (gen_tables.py).  Statically initialised __constant__ data, one copy per translation unit (no -rdc).
#include "legendre_tables.inc"
static __constant__ double kLegCoef[MLMCB200_MAX_MOMENTS + 8] = MLMCB200_LEG_COEF_INIT;   // padded: prefetch
static __constant__ double kLegAlpha[MLMCB200_MAX_MOMENTS] = MLMCB200_LEG_ALPHA_INIT;

// defined in moments.cu: acc[j] += sum_b partial[b * stride + j], j < len, fixed order (bitwise reproducible)
int launch_reduce_partials(const double* partial, int n_partials, int64_t stride, int64_t len, double* acc,
                           cudaStream_t st);

// defined in maxent.cu: the max-ent F / g / H kernels for up to 104 moments; returns 1 when the size is not covered
int64_t maxent_fast_workspace_bytes(int64_t n_nodes, int R);
int maxent_fast_launch(const double* phi, int64_t ld_g, const double* w, const double* lam, int64_t n_nodes, int R,
                       int want_h, double* out, void* workspace, int64_t workspace_bytes, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// Moments.transform (mlmc/moments.py:28-39, 58-73).  The affine map is evaluated with the same three
// IEEE operations as the reference ((v - shift) * scale + ref_lo, no FMA contraction) so that the
// closed-interval domain test makes the same decision for every finite input.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double map_to_ref(const mlmcb200_basis_t& b, double v) {
    if (b.is_log) v = log(v);
    double t = __dadd_rn(__dmul_rn(__dsub_rn(v, b.shift), b.scale), b.ref_lo);
    if (b.is_clip && (t < b.ref_lo || t > b.ref_hi)) t = __longlong_as_double(0x7ff8000000000000LL);
    return t;
}

template <bool LOG>
__device__ __forceinline__ double map_to_ref_t(const mlmcb200_basis_t& b, double v) {
    if (LOG) v = log(v);
    double t = __dadd_rn(__dmul_rn(__dsub_rn(v, b.shift), b.scale), b.ref_lo);
    if (b.is_clip && !(t >= b.ref_lo && t <= b.ref_hi)) t = __longlong_as_double(0x7ff8000000000000LL);
    return t;
}

// numpy legvander step, exactly: (p1 * t * (2i-1) - p0 * (i-1)) / i
__device__ __forceinline__ double legendre_step_exact(double p1, double p0, double t, int i) {
    return __ddiv_rn(__dsub_rn(__dmul_rn(__dmul_rn(p1, t), (double)(2 * i - 1)), __dmul_rn(p0, (double)(i - 1))),
                     (double)i);
}

// Unclipped Legendre far outside the domain: numpy's recurrence overflows to inf and then gives inf - inf = NaN two
// steps later; replay it exactly (rare path, kept out of line so that the hot kernels stay small).
static __device__ __noinline__ bool legendre_table_finite(double t, int size) {
    double p0 = 1.0, p1 = t;
    for (int i = 2; i < size; ++i) {
        const double p2 = legendre_step_exact(p1, p0, t, i);
        if (isnan(p2)) return false;
        p0 = p1;
        p1 = p2;
    }
    return true;
}

// sin and cos of a moderately sized argument (the Fourier basis maps the domain to [0, 2 pi]): Cody-Waite reduction by
// pi/2 with a two-part constant (exact for |q| < 2^20) and the fdlibm kernel polynomials on [-pi/4, pi/4]; ~25 FP64
// instructions instead of the ~70 of the general-range library routine, to which larger arguments fall back.
// Absolute error < 2e-16.
__device__ __forceinline__ void sincos_bounded(double t, double* s, double* c) {
    if (!(fabs(t) < 1.0e5)) {
        sincos(t, s, c);
        return;
    }
    const double qd = rint(t * 6.36619772367581382433e-01);                 // 2 / pi
    const int q = (int)qd;
    double r = fma(-qd, 1.57079632673412561417e+00, t);                     // pi/2, leading 33 bits
    r = fma(-qd, 6.07710050650619224932e-11, r);                            // pi/2, tail
    const double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    const double sr = fma(z * r, ps, r);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    const double cr = fma(z * z, pc, fma(z, -0.5, 1.0));
    const double sv = (q & 1) ? cr : sr, cv = (q & 1) ? sr : cr;
    *s = (q & 2) ? -sv : sv;
    *c = ((q + 1) & 2) ? -cv : cv;
}

// Does the moment vector of the mapped value t contain no NaN?  (mask_nan_samples,
// mlmc/quantity/quantity_estimate.py:6-14, applied to Moments.eval_all of this value.)
__device__ __forceinline__ bool moments_finite(const mlmcb200_basis_t& b, double t) {
    if (b.kind == MLMCB200_RAW) return !isnan(t);
    if (b.kind == MLMCB200_FOURIER) return b.size == 1 || isfinite(t);   // column 0 is the literal 1
    if (!isfinite(t)) return false;                   // numpy: v[0] = t*0 + 1 is NaN for NaN and +-inf
    if (b.kind == MLMCB200_MONOMIAL || b.is_clip || fabs(t) <= 4.0) return true;
    return legendre_table_finite(t, b.size);
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers for shared-memory staging rings ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}

__device__ __forceinline__ void mbar_init_fence() {          // make the initialised barriers visible to the async proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void bulk_copy_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done = 0;
    const uint32_t addr = smem_u32(bar);
    while (!done) {
        asm volatile(
            "{\n"
            " .reg .pred p;\n"
            " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace mlmcb200
