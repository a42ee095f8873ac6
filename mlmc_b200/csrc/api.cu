// Library-wide pieces of the C ABI: error text, device query, FP64 pipe micro-benchmarks.
#include <stdarg.h>
#include <string.h>
#include <emmintrin.h>
#include <thread>
#include <vector>
#include "common.cuh"

namespace mlmcb200 {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

int sm_count() {
    static thread_local int cached_dev = -1, cached = 0;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (dev != cached_dev) {
        int v = 0;
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
        cached = v;
        cached_dev = dev;
    }
    return cached;
}

int check_basis(const mlmcb200_basis_t* b) {
    MB_REQUIRE(b != nullptr, "null basis");
    MB_REQUIRE(b->kind >= MLMCB200_RAW && b->kind <= MLMCB200_FOURIER, "unknown basis kind %d", b->kind);
    MB_REQUIRE(b->size >= 1 && b->size <= MLMCB200_MAX_MOMENTS, "basis size %d outside [1, %d]", b->size,
               MLMCB200_MAX_MOMENTS);
    MB_REQUIRE(b->kind != MLMCB200_RAW || b->size == 1, "RAW basis must have size 1");
    return 0;
}

namespace {

// 8 independent DFMA chains per thread
__global__ void dfma_peak_kernel(double* sink, int iters, double seed) {
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6,
           a7 = seed + 7;
    const double m = 1.0000001, c = 1e-9;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
            a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
        }
    }
    const double r = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (r == 123.456) sink[0] = r;
}

// 8 independent chains, three DISTINCT register operands per DFMA (what a recurrence looks like)
__global__ void dfma3_peak_kernel(double* sink, int iters, double seed) {
    double a[8], b[8], c[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        a[i] = seed + i;
        b[i] = 1.0 + 1e-9 * (threadIdx.x + i);
        c[i] = 1e-9 * (i + 1);
    }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fma(a[i], b[i], c[i]);
#pragma unroll
            for (int i = 0; i < 8; ++i) c[i] = fma(c[i], b[i], a[i]);
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += a[i] + c[i];
    if (r == 123.456) sink[0] = r;
}

// one dependent DFMA chain per thread: cycles per dependent instruction
__global__ void dfma_latency_kernel(double* sink, int iters, double seed) {
    double a = seed;
    const double m = 1.0000001, c = 1e-9;
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 64; ++u) a = fma(a, m, c);
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        sink[0] = (double)(t1 - t0) / ((double)iters * 64.0);
        sink[1] = a;
    }
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// 8 independent 8x8 accumulator tiles per warp
__global__ void dmma_peak_kernel(double* sink, int iters, double seed) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = seed + i;
    const double a = 1.0 + 1e-9 * threadIdx.x, b = 1e-9 * (threadIdx.x & 3);
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int t = 0; t < 8; ++t) dmma884(c[2 * t], c[2 * t + 1], a, b);
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) r += c[i];
    if (r == 123.456) sink[0] = r;
}

// NACC independent accumulator tiles per warp, fragments re-read from shared memory every step (the covariance
// kernel's pattern): how many independent DMMA chains / warps does the tensor pipe need?
template <int NACC>
__global__ void dmma_chain_kernel(double* sink, int iters, double seed) {
    __shared__ double frag[64];
    if (threadIdx.x < 64) frag[threadIdx.x] = 1.0 + 1e-9 * threadIdx.x;
    __syncthreads();
    double c[2 * NACC];
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) c[i] = seed + i;
    const volatile double* fr = frag;
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 32 / NACC; ++u) {
            const double a = fr[threadIdx.x & 31], b = fr[32 + (threadIdx.x & 31)];
#pragma unroll
            for (int t = 0; t < NACC; ++t) dmma884(c[2 * t], c[2 * t + 1], a, b);
        }
    }
    double r = 0;
#pragma unroll
    for (int i = 0; i < 2 * NACC; ++i) r += c[i];
    if (r == 123.456) sink[0] = r;
}

}  // namespace
}  // namespace mlmcb200

using namespace mlmcb200;

extern "C" int mlmcb200_abi_version(void) { return MLMCB200_ABI_VERSION; }

extern "C" const char* mlmcb200_last_error(void) { return g_error; }

extern "C" int mlmcb200_sm_count(void) { return sm_count(); }

// memcpy with non-temporal stores: the staging buffer is written once and read by the DMA engine only, so the lines need
// not be fetched for ownership first (a third of the copy's DRAM traffic) nor stay in the caches.
static void copy_streaming(char* d, const char* s, size_t n) {
    if (n < 4096) {
        memcpy(d, s, n);
        return;
    }
    const size_t head = (16 - (reinterpret_cast<uintptr_t>(d) & 15)) & 15;
    if (head) {
        memcpy(d, s, head);
        d += head;
        s += head;
        n -= head;
    }
    size_t i = 0;
    for (; i + 64 <= n; i += 64) {
        const __m128i a = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i));
        const __m128i b = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 16));
        const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 32));
        const __m128i e = _mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i + 48));
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i), a);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 16), b);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 32), c);
        _mm_stream_si128(reinterpret_cast<__m128i*>(d + i + 48), e);
    }
    _mm_sfence();
    if (i < n) memcpy(d + i, s + i, n - i);
}

// The CPU stage of the staged feed (file mapping -> pinned staging buffer): n_rows rows of keep_bytes each, src_pitch
// bytes apart in the source, packed densely into dst, cut over n_threads host threads.  One thread copies a few GB/s out
// of the page cache, a tenth of the PCIe rate behind it; Python threads lose most of the gain to the interpreter lock.
extern "C" int mlmcb200_host_copy_rows(void* dst, const void* src, int64_t n_rows, int64_t keep_bytes,
                                       int64_t src_pitch, int32_t n_threads) {
    MB_REQUIRE(n_rows >= 0 && keep_bytes >= 0 && src_pitch >= keep_bytes && n_threads >= 1 && n_threads <= 256,
               "host_copy_rows: bad arguments (n_rows=%lld keep_bytes=%lld src_pitch=%lld n_threads=%d)",
               (long long)n_rows, (long long)keep_bytes, (long long)src_pitch, n_threads);
    if (n_rows == 0 || keep_bytes == 0) return 0;
    MB_REQUIRE(dst != nullptr && src != nullptr, "host_copy_rows: null buffer");
    char* const d = static_cast<char*>(dst);
    const char* const s = static_cast<const char*>(src);
    const bool dense = keep_bytes == src_pitch;
    const int64_t total = n_rows * keep_bytes;
    int64_t parts = total / (1 << 20);                               // at least 1 MB per thread
    if (parts > n_threads) parts = n_threads;
    if (!dense && parts > n_rows) parts = n_rows;
    if (parts < 1) parts = 1;
    auto work = [=](int64_t p) {
        if (dense) {                                                 // one range of bytes, cut at 64-byte multiples
            const int64_t lo = (total * p / parts) & ~int64_t(63), hi = p + 1 == parts ? total : (total * (p + 1) / parts) & ~int64_t(63);
            copy_streaming(d + lo, s + lo, (size_t)(hi - lo));
        } else {
            const int64_t r_lo = n_rows * p / parts, r_hi = n_rows * (p + 1) / parts;
            for (int64_t r = r_lo; r < r_hi; ++r) copy_streaming(d + r * keep_bytes, s + r * src_pitch, (size_t)keep_bytes);
        }
    };
    if (parts == 1) {
        work(0);
        return 0;
    }
    std::vector<std::thread> pool;
    pool.reserve((size_t)parts - 1);
    try {
        for (int64_t p = 1; p < parts; ++p) pool.emplace_back(work, p);
    } catch (...) {                                                  // could not start a thread: finish what is missing here
        const int64_t started = (int64_t)pool.size();
        work(0);
        for (int64_t p = started + 1; p < parts; ++p) work(p);
        for (auto& t : pool) t.join();
        return 0;
    }
    work(0);
    for (auto& t : pool) t.join();
    return 0;
}

extern "C" int mlmcb200_fp64_peak(int32_t kind, double* flops_per_s, void* stream) {
    // kind & 15: 0 DFMA (2 loop-invariant operands), 1 DMMA m8n8k4, 2 DFMA with 3 distinct register operands,
    //            3 dependent-DFMA latency in SM cycles (returned in *flops_per_s), 4 DMMA with (kind >> 12) & 15 = 1, 2, 4
    //            or 8 independent accumulator tiles per warp and fragments from shared memory;
    //            (kind >> 4) & 255: warps per SM (0 = 64)
    const int base = kind & 15, warps_per_sm = (kind >> 4) & 255, n_acc = (kind >> 12) & 15;
    MB_REQUIRE(flops_per_s != nullptr && base >= 0 && base <= 4, "fp64_peak: bad arguments");
    MB_REQUIRE(base != 4 || n_acc == 1 || n_acc == 2 || n_acc == 4 || n_acc == 8, "fp64_peak: bad accumulator count");
    cudaStream_t st = (cudaStream_t)stream;
    double* sink = nullptr;
    MB_CUDA_OK(cudaMalloc(&sink, 16));
    if (base == 3) {
        dfma_latency_kernel<<<1, 32, 0, st>>>(sink, 256, 0.5);
        MB_CUDA_OK(cudaGetLastError());
        MB_CUDA_OK(cudaStreamSynchronize(st));
        MB_CUDA_OK(cudaMemcpy(flops_per_s, sink, 8, cudaMemcpyDeviceToHost));
        cudaFree(sink);
        return 0;
    }
    cudaEvent_t e0, e1;
    MB_CUDA_OK(cudaEventCreate(&e0));
    MB_CUDA_OK(cudaEventCreate(&e1));
    int threads = 256, blocks = sm_count() * 8;
    if (warps_per_sm > 0) {
        threads = 128;
        blocks = sm_count() * ((warps_per_sm + 3) / 4);
    }
    const int iters = 4096;
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        MB_CUDA_OK(cudaEventRecord(e0, st));
        if (base == 0)
            dfma_peak_kernel<<<blocks, threads, 0, st>>>(sink, iters, 0.5);
        else if (base == 1)
            dmma_peak_kernel<<<blocks, threads, 0, st>>>(sink, iters, 0.5);
        else if (base == 4 && n_acc == 1)
            dmma_chain_kernel<1><<<blocks, threads, 0, st>>>(sink, iters, 0.5);
        else if (base == 4 && n_acc == 2)
            dmma_chain_kernel<2><<<blocks, threads, 0, st>>>(sink, iters, 0.5);
        else if (base == 4 && n_acc == 4)
            dmma_chain_kernel<4><<<blocks, threads, 0, st>>>(sink, iters, 0.5);
        else if (base == 4)
            dmma_chain_kernel<8><<<blocks, threads, 0, st>>>(sink, iters, 0.5);
        else
            dfma3_peak_kernel<<<blocks, threads, 0, st>>>(sink, iters, 0.5);
        MB_CUDA_OK(cudaGetLastError());
        MB_CUDA_OK(cudaEventRecord(e1, st));
        MB_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        MB_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        double flops;
        if (base == 1 || base == 4)
            flops = (double)blocks * (threads / 32) * iters * 32.0 * (8 * 8 * 4) * 2.0;  // 32 DMMA per warp-iteration
        else
            flops = (double)blocks * threads * iters * 64.0 * 2.0;                   // 64 DFMA per thread-iteration
        const double rate = flops / (ms * 1e-3);
        if (rep > 0 && rate > best) best = rate;                                        // rep 0 = warm-up
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    *flops_per_s = best;
    return 0;
}
