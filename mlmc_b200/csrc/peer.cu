// Level finalize fused with the cross-GPU reduction of the level sums, over NVLink peer memory (sm_100a).
//
// Multi-GPU estimation shards the samples of every level over the ranks; the only exchange on the
// path is the SUM of the packed level accumulators [L][2 + 2K] (2.4 kB at cfg2) before l_means / l_vars are formed
// (mlmc/quantity/quantity_estimate.py:70-77).  With NCCL that is an all-reduce launch followed by the finalize launch.
// Here ONE kernel does both: every rank stores its accumulator straight into a slot of every peer's exchange buffer
// (peer-mapped device memory, cudaIpc handles; the stores travel over NVLink / NVSwitch), raises a flag in each peer,
// waits until all flags of the current epoch have arrived in its OWN buffer, adds the P slots in rank order (so every
// rank forms bit-identical sums) and finalizes.  No host synchronisation, capturable in a CUDA graph (the epoch
// counter lives in device memory), two exchange buffers alternate by epoch parity so a fast rank cannot overwrite data
// a slow rank still reads.  Waiting is bounded (MLMCB200_PEER_TIMEOUT_MS, default 30 s, measured on %globaltimer): a rank
// that gives up POISONS the epoch in every peer's header, so a rank that arrives late fails the same epoch instead of
// returning sums its partner never saw; a failed launch leaves the local sums in `acc` untouched, writes NaN results,
// sets the sticky error word and `status[0] = 1` -- the host then repeats the reduction with NCCL on every rank
// (mlmc_b200/quantity/quantity_estimate.py).
//
// Intended for the small accumulators of scalar quantities (slot <= 64 k doubles); wide quantities keep NCCL.
#include <stdlib.h>
#include <string.h>
#include "common.cuh"

namespace mlmcb200 {
namespace {

constexpr int kPeerThreads = 256;
constexpr int kMaxWorld = 16;

struct PeerLayout {
    int64_t slot_doubles;     // capacity of one slot
    int world;
    __host__ __device__ int64_t data_offset(int buf, int slot) const {           // in doubles
        return ((int64_t)buf * world + slot) * slot_doubles;
    }
    __host__ __device__ int64_t flags_offset_bytes() const { return 2 * (int64_t)world * slot_doubles * 8; }
    __host__ __device__ int64_t header_offset_bytes() const { return flags_offset_bytes() + 2 * kMaxWorld * 4; }
    __host__ __device__ int64_t total_bytes() const { return header_offset_bytes() + 16; }
};

struct PeerArgs {
    double* acc;              // [L][acc_stride], receives the global sums
    int64_t acc_stride;
    int n_levels;
    int64_t K;
    int rank;
    PeerLayout lay;
    char* const* peers;       // device array [world]: exchange buffer of every rank as mapped in this process
    double* l_means;
    double* l_vars;
    double* mean;
    double* var;
    double* status;           // [1]: 0 = reduced, 1 = gave up (acc keeps the local sums); may be NULL
    unsigned long long timeout_ns;
};

__device__ __forceinline__ unsigned long long globaltimer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ unsigned ld_volatile_u32(const unsigned* p) { return *reinterpret_cast<const volatile unsigned*>(p); }

__global__ void __launch_bounds__(kPeerThreads) peer_allreduce_finalize_kernel(const PeerArgs a) {
    __shared__ unsigned epoch_s;
    __shared__ int failed_s;
    const int tid = threadIdx.x;
    const int world = a.lay.world;
    char* const mine = a.peers[a.rank];
    // header: [0] epoch, [1] sticky error, [2] epoch some rank gave up on (written by that rank into EVERY buffer)
    unsigned* const header = reinterpret_cast<unsigned*>(mine + a.lay.header_offset_bytes());
    if (tid == 0) {
        epoch_s = header[0] + 1u;
        failed_s = 0;
    }
    __syncthreads();
    const unsigned e = epoch_s;
    const int buf = (int)(e & 1u);
    const int64_t row = 2 + 2 * a.K;                    // doubles per level
    const int64_t n = (int64_t)a.n_levels * row;

    // 1. my sums into slot `rank` of every rank's buffer (own included)
    for (int p = 0; p < world; ++p) {
        double* dst = reinterpret_cast<double*>(a.peers[p]) + a.lay.data_offset(buf, a.rank);
        for (int64_t i = tid; i < n; i += kPeerThreads) {
            const int64_t l = i / row, j = i - l * row;
            dst[i] = a.acc[l * a.acc_stride + j];
        }
    }
    __threadfence_system();
    __syncthreads();
    // 2. raise my flag everywhere, 3. wait for everybody's flag here (bounded)
    if (tid < world) {
        unsigned* flag = reinterpret_cast<unsigned*>(a.peers[tid] + a.lay.flags_offset_bytes()) + buf * kMaxWorld + a.rank;
        *reinterpret_cast<volatile unsigned*>(flag) = e;
        const unsigned* wait_on = reinterpret_cast<const unsigned*>(mine + a.lay.flags_offset_bytes()) + buf * kMaxWorld + tid;
        const unsigned long long t0 = globaltimer_ns();
        int spins = 0;
        while (ld_volatile_u32(wait_on) != e) {
            __nanosleep(200);
            if ((++spins & 255) == 0 && globaltimer_ns() - t0 > a.timeout_ns) {
                failed_s = 1;
                break;
            }
        }
    }
    __syncthreads();
    if (failed_s != 0 && tid < world)                   // tell everybody (a late rank must fail this epoch too)
        *reinterpret_cast<volatile unsigned*>(reinterpret_cast<unsigned*>(a.peers[tid] + a.lay.header_offset_bytes()) + 2) = e;
    __threadfence_system();
    __syncthreads();
    if (tid == 0 && ld_volatile_u32(header + 2) == e) failed_s = 1;
    __syncthreads();
    const bool failed = failed_s != 0;
    // 4. global sums in rank order (identical on every rank), written back to acc
    const double* slots = reinterpret_cast<const double*>(mine) + a.lay.data_offset(buf, 0);
    for (int64_t i = tid; i < n && !failed; i += kPeerThreads) {
        double s = 0.0;
        for (int r = 0; r < world; ++r) s += __ldcv(slots + (int64_t)r * a.lay.slot_doubles + i);
        const int64_t l = i / row, j = i - l * row;
        a.acc[l * a.acc_stride + j] = s;
    }
    __syncthreads();
    // 5. finalize (quantity_estimate.py:72-77, quantity.py:592-593), same operations as finalize_levels_kernel
    for (int64_t k = tid; k < a.K; k += kPeerThreads) {
        double m_tot = 0.0, v_tot = 0.0;
        for (int l = 0; l < a.n_levels; ++l) {
            const double* r = a.acc + (int64_t)l * a.acc_stride;
            const double cnt = r[0], s = r[2 + k], sq = r[2 + a.K + k];
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            double lm = __ddiv_rn(s, cnt);
            double lv;
            if (cnt > 1.0)
                lv = __ddiv_rn(__dsub_rn(sq, __ddiv_rn(__dmul_rn(s, s), cnt)), cnt - 1.0);
            else
                lv = __longlong_as_double(0x7ff0000000000000LL);
            if (failed) lm = lv = qnan;
            if (a.l_means) a.l_means[(int64_t)l * a.K + k] = lm;
            if (a.l_vars) a.l_vars[(int64_t)l * a.K + k] = lv;
            m_tot = __dadd_rn(m_tot, lm);
            v_tot = __dadd_rn(v_tot, __ddiv_rn(lv, cnt));
        }
        if (a.mean) a.mean[k] = m_tot;
        if (a.var) a.var[k] = v_tot;
    }
    if (tid == 0) {
        header[0] = e;
        if (failed) header[1] = 1u;
        if (a.status) a.status[0] = failed ? 1.0 : 0.0;
    }
}

unsigned long long peer_timeout_ns() {
    static unsigned long long cached = 0;
    if (cached == 0) {
        const char* e = getenv("MLMCB200_PEER_TIMEOUT_MS");
        const double ms = e ? atof(e) : 30000.0;
        cached = (unsigned long long)((ms > 1.0 ? ms : 1.0) * 1e6);
    }
    return cached;
}

}  // namespace
}  // namespace mlmcb200

using namespace mlmcb200;

extern "C" int64_t mlmcb200_peer_buffer_bytes(int32_t world, int64_t slot_doubles) {
    if (world < 1 || world > kMaxWorld || slot_doubles < 1) return -1;
    PeerLayout lay{slot_doubles, world};
    return lay.total_bytes();
}

extern "C" int mlmcb200_peer_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64) {
    MB_REQUIRE(bytes > 0 && dev_ptr != nullptr && handle64 != nullptr, "peer_alloc: bad arguments");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
    void* p = nullptr;
    MB_CUDA_OK(cudaMalloc(&p, (size_t)bytes));
    MB_CUDA_OK(cudaMemset(p, 0, (size_t)bytes));
    cudaIpcMemHandle_t h;
    MB_CUDA_OK(cudaIpcGetMemHandle(&h, p));
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
    return 0;
}

extern "C" int mlmcb200_peer_open(const unsigned char* handle64, void** dev_ptr) {
    MB_REQUIRE(handle64 != nullptr && dev_ptr != nullptr, "peer_open: bad arguments");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    MB_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
    return 0;
}

extern "C" int mlmcb200_peer_close(void* dev_ptr) {
    if (dev_ptr != nullptr) MB_CUDA_OK(cudaIpcCloseMemHandle(dev_ptr));
    return 0;
}

extern "C" int mlmcb200_peer_free(void* dev_ptr) {
    if (dev_ptr != nullptr) MB_CUDA_OK(cudaFree(dev_ptr));
    return 0;
}

extern "C" int mlmcb200_peer_error(const void* own_buffer, int32_t world, int64_t slot_doubles, int32_t* error) {
    MB_REQUIRE(own_buffer != nullptr && error != nullptr && world >= 1 && world <= kMaxWorld, "peer_error: bad arguments");
    PeerLayout lay{slot_doubles, world};
    unsigned hdr[2] = {0, 0};
    MB_CUDA_OK(cudaMemcpy(hdr, static_cast<const char*>(own_buffer) + lay.header_offset_bytes(), 8, cudaMemcpyDeviceToHost));
    *error = (int32_t)hdr[1];
    return 0;
}

extern "C" int mlmcb200_allreduce_finalize_levels(double* acc, int64_t acc_stride, int32_t n_levels, int64_t K,
                                                  int32_t rank, int32_t world, void* const* peer_buffers,
                                                  int64_t slot_doubles, double* l_means, double* l_vars, double* mean,
                                                  double* var, double* status, void* stream) {
    MB_REQUIRE(acc != nullptr && peer_buffers != nullptr && n_levels >= 1 && K >= 1 && acc_stride >= 2 + 2 * K,
               "allreduce_finalize_levels: bad arguments");
    MB_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "allreduce_finalize_levels: rank %d of %d",
               rank, world);
    MB_REQUIRE((int64_t)n_levels * (2 + 2 * K) <= slot_doubles, "allreduce_finalize_levels: %lld doubles exceed the slot (%lld)",
               (long long)n_levels * (2 + 2 * K), (long long)slot_doubles);
    PeerArgs a;
    a.acc = acc;
    a.acc_stride = acc_stride;
    a.n_levels = n_levels;
    a.K = K;
    a.rank = rank;
    a.lay = PeerLayout{slot_doubles, world};
    a.peers = reinterpret_cast<char* const*>(peer_buffers);
    a.l_means = l_means;
    a.l_vars = l_vars;
    a.mean = mean;
    a.var = var;
    a.status = status;
    a.timeout_ns = peer_timeout_ns();
    peer_allreduce_finalize_kernel<<<1, kPeerThreads, 0, (cudaStream_t)stream>>>(a);
    MB_CUDA_OK(cudaGetLastError());
    return 0;
}
