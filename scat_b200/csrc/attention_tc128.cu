// Softmax attention for n = 128 tokens x 64-wide heads on tcgen05 (the HRNet-token variant, BASELINE config 4:
// hand_net.py:161,193-203 with vision_transformer.py:61-77).  At n = 128 a (batch, head) problem IS one UMMA tile:
//
//   S = Q K^T      128 x 128 x 64   tcgen05.mma kind::tf32, A = Q and B = K both K-major, accumulator in TMEM
//   P = softmax(S * 64^-0.5)        tcgen05.ld hands every thread one full row of S: max / exp / sum in registers
//   O = P V        128 x 64 x 128   A = P (written back to shared memory in the K-major 128B-swizzle layout),
//                                   B = V^T read MN-major straight from the V columns of the QKV matrix
//
// One CTA per problem; warp 0 issues the TMA loads (Q, K, V tiles cut directly out of the [B*128, 3*inner] QKV
// matrix by two tensor maps) and the MMAs, warps 1-4 own the four 32-lane TMEM quadrants for softmax and the
// epilogue.  P aliases the Q/K staging buffers (dead once S is complete), so a CTA needs 96 KB of shared memory and
// 256 TMEM columns: two CTAs per SM overlap each other's load, softmax and store phases.  fp32 operands are
// truncated to TF32 by the tensor core; P is rounded to TF32-nearest when it is written.  Inference only (the
// reference never trains this variant: SURVEY.md section 0), so P is not saved.
#include <cuda.h>

#include "kernels.h"

namespace scat {
namespace {

constexpr int N = 128, DH = 64;
constexpr int THREADS = 160;
constexpr uint32_t SPIN_LIMIT = 1u << 28;
constexpr int QK_BYTES = N * DH * 4;                 // 32 KB each, as 2 k-blocks of [128 rows x 128 B]
constexpr int V_BYTES = N * DH * 4;                  // 32 KB, as [4 k-blocks][2 n-boxes][32 k x 128 B]
constexpr int OFF_Q = 0, OFF_K = QK_BYTES, OFF_V = 2 * QK_BYTES, OFF_P = 0;   // P [128 x 128] fp32 = 64 KB over Q | K
constexpr int OFF_BAR = 3 * QK_BYTES;
constexpr int SMEM_TOTAL = OFF_BAR + 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {     // bounded: a protocol bug traps, never hangs
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity))
        if (++spins > SPIN_LIMIT) __trap();
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint32_t bar, uint32_t dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc,
                                          uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
                 "setp.ne.b32 p, %6, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}"
                 ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(THREADS, 2)
attention_fwd_tc128_kernel(const __grid_constant__ CUtensorMap tmK /* K-major boxes {32 d, 128 rows} */,
                           const __grid_constant__ CUtensorMap tmV /* MN-major boxes {32 d, 32 rows} */,
                           float* __restrict__ O, int heads, int out_mode) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);
    const uint32_t bar_qk = sb + OFF_BAR, bar_v = bar_qk + 8, bar_s = bar_qk + 16, bar_p = bar_qk + 24, bar_o = bar_qk + 32;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + OFF_BAR + 40);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    const int inner = heads * DH;
    if (threadIdx.x == 0 && (sb & 1023u) != 0) __trap();

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmK)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmV)) : "memory");
        mbar_init(bar_qk, 1); mbar_init(bar_v, 1); mbar_init(bar_s, 1); mbar_init(bar_p, 128); mbar_init(bar_o, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_S = *tmem_slot, tmem_O = tmem_S + 128;
    pdl_sync();

    // descriptor constants (see gemm_tc.cu): K-major 128B swizzle: SBO 1024, layout 2, k-step +32 B;
    // MN-major tf32 (SWIZZLE_128B_BASE32B): LBO = one 4 KB box, SBO 512, layout 1, k-step +1024 B
    constexpr uint32_t k_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    constexpr uint32_t mn_hi = (512u >> 4) | (1u << 14) | (1u << 29);
    constexpr uint32_t idesc_s = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    constexpr uint32_t idesc_o = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 16) | ((uint32_t)(DH >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);

    if (warp == 0) {
        if (elect_one()) {
            const int row0 = b * N;
            mbar_expect_tx(bar_qk, 2 * QK_BYTES);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb) {
                tma_load_2d(&tmK, bar_qk, sb + OFF_Q + kb * 16384, h * DH + kb * 32, row0);             // Q columns
                tma_load_2d(&tmK, bar_qk, sb + OFF_K + kb * 16384, inner + h * DH + kb * 32, row0);     // K columns
            }
            mbar_expect_tx(bar_v, V_BYTES);
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int nb = 0; nb < 2; ++nb)
                    tma_load_2d(&tmV, bar_v, sb + OFF_V + (kb * 2 + nb) * 4096, 2 * inner + h * DH + nb * 32, row0 + kb * 32);
        }
        __syncwarp();
        // S = Q K^T
        mbar_wait(bar_qk, 0);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t q_lo = ((sb + OFF_Q) >> 4) | (1u << 16), kk_lo = ((sb + OFF_K) >> 4) | (1u << 16);
#pragma unroll
            for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_tf32(tmem_S, q_lo + kb * (16384 >> 4) + k * 2, k_hi, kk_lo + kb * (16384 >> 4) + k * 2, k_hi, idesc_s,
                              (kb | k) != 0 ? 1u : 0u);
            umma_commit(bar_s);
        }
        __syncwarp();
        // O = P V once the softmax warps have written P (over the Q/K buffers) and V has landed
        mbar_wait(bar_p, 0);
        mbar_wait(bar_v, 0);
        tc_fence_after();
        if (elect_one()) {
            const uint32_t p_lo = ((sb + OFF_P) >> 4) | (1u << 16);
            const uint32_t v_lo = ((sb + OFF_V) >> 4) | ((4096u >> 4) << 16);
#pragma unroll
            for (int kb = 0; kb < 4; ++kb)
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_tf32(tmem_O, p_lo + kb * (16384 >> 4) + k * 2, k_hi, v_lo + kb * (8192 >> 4) + k * (1024 >> 4), mn_hi,
                              idesc_o, (kb | k) != 0 ? 1u : 0u);
            umma_commit(bar_o);
        }
        __syncwarp();
    } else {
        const int q = warp & 3;                         // TMEM lane quadrant of this warp
        const int r = q * 32 + lane;                    // query row i of this thread
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        mbar_wait(bar_s, 0);
        tc_fence_after();
        // row-wise softmax of S * 64^-0.5: the whole row lives in this thread (4 TMEM loads of 32 columns)
        float m = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < N; c += 32) {
            float v[32];
            tmem_ld32(tmem_S + lane_base + c, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) m = fmaxf(m, v[j]);
        }
        m *= 0.125f;
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < N; c += 32) {
            float v[32];
            tmem_ld32(tmem_S + lane_base + c, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) sum += expf(v[j] * 0.125f - m);
        }
        const float inv = 1.0f / sum;
        // P -> shared memory, K-major 128B-swizzle layout of the A operand: k-block kb = c / 32 is a [128 x 128 B] slab,
        // 8-row groups 1024 B apart, 16-byte chunk ch of row r stored at chunk ch ^ (r & 7)
        uint8_t* prow = smem + OFF_P + (r >> 3) * 1024 + (r & 7) * 128;
#pragma unroll 1
        for (int c = 0; c < N; c += 32) {
            float v[32];
            tmem_ld32(tmem_S + lane_base + c, v);
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                float4 p4;
                p4.x = round_tf32(expf(v[4 * ch] * 0.125f - m) * inv);
                p4.y = round_tf32(expf(v[4 * ch + 1] * 0.125f - m) * inv);
                p4.z = round_tf32(expf(v[4 * ch + 2] * 0.125f - m) * inv);
                p4.w = round_tf32(expf(v[4 * ch + 3] * 0.125f - m) * inv);
                *reinterpret_cast<float4*>(prow + (c >> 5) * 16384 + ((ch ^ (r & 7)) << 4)) = p4;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // generic-proxy writes -> UMMA (async proxy) reads
        tc_fence_before();                                                   // the TMEM reads of S are done
        mbar_arrive(bar_p);
        // epilogue: O row -> global (head-merged [B*n, inner], vision_transformer.py:77)
        mbar_wait(bar_o, 0);
        tc_fence_after();
        const long long orow = ((long long)b * N + r) * inner + h * DH;
#pragma unroll 1
        for (int c = 0; c < DH; c += 32) {
            float v[32];
            tmem_ld32(tmem_O + lane_base + c, v);
#pragma unroll
            for (int j = 0; j < 32; j += 4) store_out4(O, orow + c + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]), out_mode);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_S), "r"(256u));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(ptr);
    }
    return fn;
}

}  // namespace

bool attention_tc128_supported(int n, int heads, const float* QKV) {
    return n == N && heads >= 1 && (reinterpret_cast<uintptr_t>(QKV) & 15) == 0 && encode_fn() != nullptr;
}

int launch_attention_tc128_fwd(const float* QKV, float* O, int B, int n, int heads, int out_mode, cudaStream_t stream) {
    SCAT_REQUIRE(n == N, kErrUnsupported, "attention_tc128: n=%d", n);
    EncodeTiledFn fn = encode_fn();
    SCAT_REQUIRE(fn != nullptr, kErrUnsupported, "cuTensorMapEncodeTiled entry point not available");
    const int inner = heads * DH;
    cuuint64_t dims[2] = {(cuuint64_t)(3 * inner), (cuuint64_t)B * N};
    cuuint64_t strides[1] = {(cuuint64_t)(3 * inner) * sizeof(float)};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMap tmK, tmV;
    cuuint32_t boxK[2] = {32, 128}, boxV[2] = {32, 32};
    CUresult r1 = fn(&tmK, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(QKV), dims, strides, boxK, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CUresult r2 = fn(&tmV, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(QKV), dims, strides, boxV, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    SCAT_REQUIRE(r1 == CUDA_SUCCESS && r2 == CUDA_SUCCESS, kErrUnsupported, "attention_tc128: tensor map encode failed (%d %d)",
                 (int)r1, (int)r2);
    SCAT_ENSURE_SMEM(attention_fwd_tc128_kernel, SMEM_TOTAL);
    SCAT_CHECK_CUDA(launch_k(attention_fwd_tc128_kernel, dim3(B * heads), dim3(THREADS), SMEM_TOTAL, stream, tmK, tmV, O, heads, out_mode));
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace scat
