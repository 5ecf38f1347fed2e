// Warp-per-(sample, head) attention for short sequences (n <= 32; the hand head has n = 21 tokens).
// vision_transformer.py:61-77:  S = Q K^T * 64^-0.5,  P = softmax(S),  O = P V   and its backward.
//
// One warp owns one (b, h) problem; there are no block-wide barriers, only __syncwarp.  In the score phase
// lane j keeps row j of K (or V) in registers and streams rows of Q (or dO) from shared memory as broadcast
// float4 loads, so the inner loops are 64 FMAs per 16 shared-memory loads; softmax reductions over j are warp
// shuffles; in the output phase lane i accumulates row i of O (or dQ) in registers.
#include "kernels.h"

namespace scat {
namespace {

constexpr int DH = 64;
constexpr int WPB = 4;   // warps (problems) per CTA

__device__ __forceinline__ void load_rows_to_smem(const float* __restrict__ src, long long row_stride, int n, float* dst,
                                                  int lane) {
    for (int e = lane; e < n * (DH / 4); e += 32) {
        const int r = e >> 4, c4 = e & 15;
        reinterpret_cast<float4*>(dst)[e] = __ldg(reinterpret_cast<const float4*>(src + (long long)r * row_stride) + c4);
    }
}
__device__ __forceinline__ void load_row_to_regs(const float* __restrict__ src, float* reg) {
#pragma unroll
    for (int c4 = 0; c4 < DH / 4; ++c4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + c4);
        reg[4 * c4] = v.x; reg[4 * c4 + 1] = v.y; reg[4 * c4 + 2] = v.z; reg[4 * c4 + 3] = v.w;
    }
}
__device__ __forceinline__ float dot_row(const float* srow, const float* reg) {
    float a = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < DH / 4; ++c4) {
        const float4 q = reinterpret_cast<const float4*>(srow)[c4];
        a = fmaf(q.x, reg[4 * c4], a); a = fmaf(q.y, reg[4 * c4 + 1], a);
        a = fmaf(q.z, reg[4 * c4 + 2], a); a = fmaf(q.w, reg[4 * c4 + 3], a);
    }
    return a;
}
__device__ __forceinline__ void axpy_row(float w, const float* srow, float* acc) {
#pragma unroll
    for (int c4 = 0; c4 < DH / 4; ++c4) {
        const float4 v = reinterpret_cast<const float4*>(srow)[c4];
        acc[4 * c4] = fmaf(w, v.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(w, v.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(w, v.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(w, v.w, acc[4 * c4 + 3]);
    }
}
__device__ __forceinline__ void store_row(float* dst, const float* acc, int round_out) {
#pragma unroll
    for (int c4 = 0; c4 < DH / 4; ++c4) {
        float4 v = make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]);
        if (round_out) { v.x = round_tf32(v.x); v.y = round_tf32(v.y); v.z = round_tf32(v.z); v.w = round_tf32(v.w); }
        reinterpret_cast<float4*>(dst)[c4] = v;
    }
}

template <int N>
__global__ void __launch_bounds__(WPB * 32)
attention_fwd_warp_kernel(const float* __restrict__ QKV, float* __restrict__ O, float* __restrict__ P, int heads,
                          int problems, int round_out) {
    extern __shared__ __align__(16) float sm[];
    constexpr int PER_WARP = (2 * N * DH + N * (N + 1) + 3) / 4 * 4;     // keep every warp's slab float4-aligned
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int prob = blockIdx.x * WPB + warp;
    if (prob >= problems) return;
    float* Qs = sm + warp * PER_WARP;
    float* Vs = Qs + N * DH;
    float* Ps = Vs + N * DH;
    const int b = prob / heads, g = prob % heads;
    const int inner = heads * DH;
    const long long rs = 3LL * inner;
    const float* base = QKV + (long long)b * N * rs + g * DH;
    load_rows_to_smem(base, rs, N, Qs, lane);
    load_rows_to_smem(base + 2 * inner, rs, N, Vs, lane);
    float kreg[DH];
    const bool act = lane < N;
    if (act) load_row_to_regs(base + inner + (long long)lane * rs, kreg);
    __syncwarp();
    float s[N];
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = act ? dot_row(Qs + i * DH, kreg) * 0.125f : -INFINITY;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float m = warp_max(s[i]);
        const float e = act ? expf(s[i] - m) : 0.f;
        const float sum = warp_sum(e);
        s[i] = e * (1.0f / sum);
        if (act) Ps[i * (N + 1) + lane] = s[i];
    }
    __syncwarp();
    float* Pg = P + (long long)prob * N * N;
    for (int e = lane; e < N * N; e += 32) Pg[e] = Ps[(e / N) * (N + 1) + (e % N)];
    if (act) {
        float acc[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] = 0.f;
#pragma unroll 3
        for (int j = 0; j < N; ++j) axpy_row(Ps[lane * (N + 1) + j], Vs + j * DH, acc);
        store_row(O + ((long long)b * N + lane) * inner + g * DH, acc, round_out);
    }
}

template <int N>
__global__ void __launch_bounds__(WPB * 32)
attention_bwd_warp_kernel(const float* __restrict__ QKV, const float* __restrict__ P, const float* __restrict__ dO,
                          float* __restrict__ dQKV, int heads, int problems, int round_out) {
    extern __shared__ __align__(16) float sm[];
    constexpr int PER_WARP = (3 * N * DH + N * (N + 1) + 3) / 4 * 4;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int prob = blockIdx.x * WPB + warp;
    if (prob >= problems) return;
    float* Qs = sm + warp * PER_WARP;
    float* Ks = Qs + N * DH;
    float* Gs = Ks + N * DH;
    float* Ss = Gs + N * DH;
    const int b = prob / heads, g = prob % heads;
    const int inner = heads * DH;
    const long long rs = 3LL * inner;
    const float* base = QKV + (long long)b * N * rs + g * DH;
    load_rows_to_smem(base, rs, N, Qs, lane);
    load_rows_to_smem(base + inner, rs, N, Ks, lane);
    load_rows_to_smem(dO + (long long)b * N * inner + g * DH, inner, N, Gs, lane);
    const bool act = lane < N;
    const float* Pg = P + (long long)prob * N * N;
    float p[N], ds[N];
    {
        float vreg[DH];
        if (act) load_row_to_regs(base + 2 * inner + (long long)lane * rs, vreg);
#pragma unroll
        for (int i = 0; i < N; ++i) p[i] = act ? __ldg(Pg + i * N + lane) : 0.f;
        __syncwarp();
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const float dp = act ? dot_row(Gs + i * DH, vreg) : 0.f;      // dP[i,j] = dO[i] . V[j]
            const float r = warp_sum(dp * p[i]);                          // rowsum(dP * P)
            ds[i] = p[i] * (dp - r) * 0.125f;                             // dS (scale folded in)
            if (act) Ss[i * (N + 1) + lane] = ds[i];
        }
    }
    float* dbase = dQKV + (long long)b * N * rs + g * DH;
    if (act) {
        float acc[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] = 0.f;
#pragma unroll 3
        for (int i = 0; i < N; ++i) axpy_row(p[i], Gs + i * DH, acc);     // dV[j] = sum_i P[i,j] dO[i]
        store_row(dbase + 2 * inner + (long long)lane * rs, acc, round_out);
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] = 0.f;
#pragma unroll 3
        for (int i = 0; i < N; ++i) axpy_row(ds[i], Qs + i * DH, acc);    // dK[j] = sum_i dS[i,j] Q[i]
        store_row(dbase + inner + (long long)lane * rs, acc, round_out);
    }
    __syncwarp();
    if (act) {
        float acc[DH];
#pragma unroll
        for (int d = 0; d < DH; ++d) acc[d] = 0.f;
#pragma unroll 3
        for (int j = 0; j < N; ++j) axpy_row(Ss[lane * (N + 1) + j], Ks + j * DH, acc);   // dQ[i] = sum_j dS[i,j] K[j]
        store_row(dbase + (long long)lane * rs, acc, round_out);
    }
}

template <int N>
int launch_fwd(const float* QKV, float* O, float* P, int B, int heads, int round_out, cudaStream_t stream) {
    const size_t smem = sizeof(float) * WPB * ((2 * N * DH + N * (N + 1) + 3) / 4 * 4);
    static bool done = false;
    if (!done) {
        SCAT_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_warp_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done = true;
    }
    const int problems = B * heads;
    attention_fwd_warp_kernel<N><<<ceil_div(problems, WPB), WPB * 32, smem, stream>>>(QKV, O, P, heads, problems, round_out);
    SCAT_CHECK_LAUNCH();
    return 0;
}
template <int N>
int launch_bwd(const float* QKV, const float* P, const float* dO, float* dQKV, int B, int heads, int round_out,
               cudaStream_t stream) {
    const size_t smem = sizeof(float) * WPB * ((3 * N * DH + N * (N + 1) + 3) / 4 * 4);
    static bool done = false;
    if (!done) {
        SCAT_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_warp_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        done = true;
    }
    const int problems = B * heads;
    attention_bwd_warp_kernel<N><<<ceil_div(problems, WPB), WPB * 32, smem, stream>>>(QKV, P, dO, dQKV, heads, problems, round_out);
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace

// n == 21 (one token per joint, hand_net.py:328) is the only short length on the reference path
bool attention_small_supported(int n) { return n == 21; }
int launch_attention_small_fwd(const float* QKV, float* O, float* P, int B, int n, int heads, int round_out,
                               cudaStream_t stream) {
    SCAT_REQUIRE(n == 21, kErrUnsupported, "attention_small: n=%d", n);
    return launch_fwd<21>(QKV, O, P, B, heads, round_out, stream);
}
int launch_attention_small_bwd(const float* QKV, const float* P, const float* dO, float* dQKV, int B, int n, int heads,
                               int round_out, cudaStream_t stream) {
    SCAT_REQUIRE(n == 21, kErrUnsupported, "attention_small: n=%d", n);
    return launch_bwd<21>(QKV, P, dO, dQKV, B, heads, round_out, stream);
}

}  // namespace scat
