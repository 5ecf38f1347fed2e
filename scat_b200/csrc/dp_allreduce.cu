// Gradient all-reduce over NVLink peer memory for the data-parallel train step (the reference
// has no distributed code -- DistributedDataParallel is imported at train.py:18 and never used).
//
// One process per GPU.  Each rank owns a gradient bucket and a small signal area, both plain cudaMalloc memory
// exported with CUDA IPC and mapped by every peer of the node.  One kernel per step does the whole exchange
// ("two shot"): rank r sums slice r of all buckets with 128-bit loads over NVLink, in rank order (so every rank holds
// bit-identical sums), and stores the result into slice r of every bucket.  Per GPU that is (W-1)/W of the bucket in
// and out, the minimum for an all-reduce, and no staging copy.  Cross-GPU ordering is a flag barrier per thread
// block at both ends: block b of every rank raises a flag in every peer's signal area (st.release.sys) and waits for
// all peers' flags in its own (ld.acquire.sys).  The first barrier says "my gradients are final and I no longer read
// yours from the previous step"; the second says "all my stores into your bucket have landed".  The kernel is an
// ordinary launch on the step's stream, so it is captured into the step's CUDA graph and costs no host work.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace scat {
namespace {

constexpr int kMaxPeers = 8;
constexpr int kMaxBlocks = 296;                 // size of the per-block flag arrays in the signal area
// The exchange runs BESIDE the backward's kernels (HeadTrainStep hides it there), so it has to fit next to them.  A
// tensor-core GEMM CTA holds 32 K of an SM's 64 K registers (two CTAs per SM), a persistent conv CTA 41 K: wherever a
// block of this kernel sits, one GEMM CTA does not fit until it leaves -- and it spends most of its life waiting for NVLink
// round trips and for the peers.  Measured at 2 GPUs (profiles/r2_exchange_probe.txt): 296 x 512 threads (round 1, the
// whole register file of every SM) made conv wgrad 22 -> 46 us and the GEMM beside it 18 -> 35 us; 148 x 256 still
// touched every SM.  So: few blocks (that many SMs lose one CTA slot, the rest of the GPU does not notice), 256 threads x
// up to 4 x 16 B of loads in flight each.  Block-count sweep at 2 GPUs, step with the exchange hidden / exposed after the
// last kernel: 32 blocks 0.796 / 0.840 ms, 64: 0.805 / 0.826, 148: 0.802 / 0.829 (no exchange: 0.775) -- the default serves
// the hidden form.
// SCAT_PEER_BLOCKS overrides the block count (every rank must use the same value).
// (defaults per world size in launch_w: 48 / 64 / 128 blocks for 2 / 4 / 8 ranks -- the 9 MB layer-0 part has ~70 us of conv
// backward to hide under, and more ranks mean longer round trips for the same bytes)
constexpr int kThreads = 256;
// signal area (uint32 words): flags[kMaxBlocks][kMaxPeers], epoch[kMaxBlocks], error
constexpr int kSigFlags = 0;
constexpr int kSigEpoch = kMaxBlocks * kMaxPeers;
constexpr int kSigError = kSigEpoch + kMaxBlocks;
constexpr int kSigWords = kSigError + 8;
constexpr unsigned long long kSpinLimitNs = 20ull * 1000 * 1000 * 1000;   // a missing peer must not hang the GPU

struct PeerSet {
    float* data[kMaxPeers];
    uint32_t* sig[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {
    float4 v;
    asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// all ranks' block `blockIdx.x` meet here; `val` increases by one per barrier.  The bar.sync orders every thread's
// earlier stores before the flag stores of threads 0..W-1, and st.release.sys is cumulative over them, so no
// per-thread system fence is needed.  Returns false (for the whole block) when a peer did not arrive within the
// bounded wait: the error word of this rank's signal area is then set and STAYS set.
// Forward progress: block b only waits for block b of the peers, every rank launches the same grid and the hardware
// dispatches blocks in ascending index order, so the lowest unfinished block index is resident on every rank.
template <int W>
__device__ __forceinline__ bool peer_barrier(const PeerSet& ps, int rank, uint32_t val) {
    __shared__ int failed;
    if (threadIdx.x == 0) failed = 0;
    __syncthreads();
    if (threadIdx.x < W) {
        st_release_sys(ps.sig[threadIdx.x] + kSigFlags + blockIdx.x * kMaxPeers + rank, val);
        const uint32_t* mine = ps.sig[rank] + kSigFlags + blockIdx.x * kMaxPeers + threadIdx.x;
        unsigned long long t0 = 0;
        uint32_t spins = 0;
        while ((int32_t)(ld_acquire_sys(mine) - val) < 0) {
            if ((++spins & 0xfffu) == 0) {               // look at the clock every 4096 polls only
                const unsigned long long now = global_ns();
                if (t0 == 0) t0 = now;
                if (now - t0 > kSpinLimitNs) {
                    ps.sig[rank][kSigError] = 1u;
                    failed = 1;
                    break;
                }
            }
        }
    }
    __syncthreads();
    return failed == 0;
}

template <int W>
__global__ void __launch_bounds__(kThreads) peer_allreduce_kernel(PeerSet ps, int rank, long long lo4, long long n4, int last) {
    // loads of U grid-strides are issued together: an NVLink round trip is several microseconds, and a volatile load
    // is not moved across the stores of the previous iteration
    constexpr int U = W >= 8 ? 1 : (W == 4 ? 2 : 4);
    pdl_sync();                                  // this rank's gradients are final from here on
    // a wait that timed out is sticky: this and every later exchange on this rank does nothing (the partial sums of a
    // broken exchange must never reach the weights; the optimiser kernel reads the same word and skips its update)
    if (*reinterpret_cast<volatile uint32_t*>(ps.sig[rank] + kSigError) != 0u) return;
    uint32_t* epoch = ps.sig[rank] + kSigEpoch + blockIdx.x;
    const uint32_t e = *epoch;
    if (!peer_barrier<W>(ps, rank, 2 * e + 1)) return;
    const long long per = (n4 + W - 1) / W;
    const long long begin = lo4 + (long long)rank * per;
    const long long end = min(begin + per, lo4 + n4);
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long i0 = begin + (long long)blockIdx.x * kThreads + threadIdx.x; i0 < end; i0 += stride * U) {
        float4 v[U][W];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            if (i < end) {
#pragma unroll
                for (int p = 0; p < W; ++p) v[u][p] = ld_peer(reinterpret_cast<const float4*>(ps.data[p]) + i);
            } else {
#pragma unroll
                for (int p = 0; p < W; ++p) v[u][p] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long long i = i0 + u * stride;
            float4 s = v[u][0];
#pragma unroll
            for (int p = 1; p < W; ++p) {
                s.x += v[u][p].x; s.y += v[u][p].y; s.z += v[u][p].z; s.w += v[u][p].w;
            }
            if (i < end) {
#pragma unroll
                for (int p = 0; p < W; ++p) reinterpret_cast<float4*>(ps.data[p])[i] = s;
            }
        }
    }
    // the closing barrier (every peer has stored its sums into this rank's bucket) may be left to a LATER exchange of the
    // same stream: its opening barrier is reached by a rank only after that rank's earlier exchanges have stored, and
    // nothing reads the sums or rewrites the gradients before the last exchange of the step has closed
    if (last) peer_barrier<W>(ps, rank, 2 * e + 2);
    if (threadIdx.x == 0) *epoch = e + 1;
}

template <int W>
int launch_w(const PeerSet& ps, int rank, long long lo4, long long n4, int last, cudaStream_t st) {
    constexpr int U = W >= 8 ? 1 : (W == 4 ? 2 : 4);        // as in the kernel
    const long long per = (n4 + W - 1) / W, per_block = (long long)kThreads * U;
    static const int env_blocks = [] { const char* e = getenv("SCAT_PEER_BLOCKS"); return e ? atoi(e) : 0; }();
    // more ranks = more NVLink round trips per byte of this rank's share: the block count grows with the world size
    const int launch_blocks = std::max(1, std::min(env_blocks > 0 ? env_blocks : (W >= 8 ? 128 : W >= 4 ? 64 : 48), kMaxBlocks));
    const int grid = (int)std::min<long long>(launch_blocks, std::max<long long>(1, (per + per_block - 1) / per_block));
    SCAT_CHECK_CUDA(launch_k(peer_allreduce_kernel<W>, dim3(grid), dim3(kThreads), 0, st, ps, rank, lo4, n4, last));
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace
}  // namespace scat

using namespace scat;

extern "C" {

size_t scat_peer_signal_bytes(void) { return kSigWords * sizeof(uint32_t); }

int scat_peer_alloc(size_t bytes, void** out) {
    SCAT_REQUIRE(out && bytes > 0, kErrBadArg, "peer_alloc: bad argument");
    SCAT_CHECK_CUDA(cudaMalloc(out, bytes));
    SCAT_CHECK_CUDA(cudaMemset(*out, 0, bytes));
    SCAT_CHECK_CUDA(cudaDeviceSynchronize());
    return 0;
}

int scat_peer_free(void* ptr) {
    SCAT_CHECK_CUDA(cudaFree(ptr));
    return 0;
}

int scat_peer_export(void* ptr, uint8_t* handle64) {
    SCAT_REQUIRE(ptr && handle64, kErrBadArg, "peer_export: null");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t h;
    SCAT_CHECK_CUDA(cudaIpcGetMemHandle(&h, ptr));
    memcpy(handle64, &h, 64);
    return 0;
}

int scat_peer_open(const uint8_t* handle64, void** out) {
    SCAT_REQUIRE(handle64 && out, kErrBadArg, "peer_open: null");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    SCAT_CHECK_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess));
    return 0;
}

int scat_peer_close(void* ptr) {
    SCAT_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
    return 0;
}

static int peer_allreduce_impl(float* const* buckets, uint32_t* const* signals, int32_t rank, int32_t world, long long lo,
                               long long hi, int last, void* stream) {
    SCAT_REQUIRE(buckets && signals, kErrBadArg, "peer_allreduce: null");
    SCAT_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world, kErrBadArg,
                 "peer_allreduce: rank %d of %d (at most %d peers)", rank, world, kMaxPeers);
    SCAT_REQUIRE(lo >= 0 && hi > lo && lo % 4 == 0 && hi % 4 == 0, kErrBadArg,
                 "peer_allreduce: range [%lld, %lld) must be non-empty and 16-byte aligned", lo, hi);
    PeerSet ps = {};
    for (int p = 0; p < world; ++p) {
        SCAT_REQUIRE(buckets[p] && signals[p], kErrBadArg, "peer_allreduce: peer %d not mapped", p);
        ps.data[p] = buckets[p];
        ps.sig[p] = signals[p];
    }
    cudaStream_t st = (cudaStream_t)stream;
    const long long lo4 = lo / 4, n4 = (hi - lo) / 4;
    switch (world) {
        case 1: return launch_w<1>(ps, rank, lo4, n4, last, st);
        case 2: return launch_w<2>(ps, rank, lo4, n4, last, st);
        case 4: return launch_w<4>(ps, rank, lo4, n4, last, st);
        case 8: return launch_w<8>(ps, rank, lo4, n4, last, st);
        default: break;
    }
    SCAT_REQUIRE(false, kErrUnsupported, "peer_allreduce: world size %d (1, 2, 4 or 8)", world);
}

int scat_peer_allreduce(float* const* buckets, uint32_t* const* signals, int32_t rank, int32_t world, long long lo,
                        long long hi, void* stream) {
    return peer_allreduce_impl(buckets, signals, rank, world, lo, hi, 1, stream);
}

int scat_peer_allreduce_part(float* const* buckets, uint32_t* const* signals, int32_t rank, int32_t world, long long lo,
                             long long hi, int32_t last, void* stream) {
    return peer_allreduce_impl(buckets, signals, rank, world, lo, hi, last ? 1 : 0, stream);
}

const uint32_t* scat_peer_error_word(const uint32_t* signal) { return signal ? signal + kSigError : nullptr; }

int scat_peer_error(const uint32_t* signal, int32_t* out) {
    SCAT_REQUIRE(signal && out, kErrBadArg, "peer_error: null");
    uint32_t v = 0;
    SCAT_CHECK_CUDA(cudaMemcpy(&v, signal + kSigError, sizeof(v), cudaMemcpyDeviceToHost));
    *out = (int32_t)v;
    return 0;
}

}  // extern "C"
