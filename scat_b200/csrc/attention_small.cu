// CTA-per-(sample, head) attention for short sequences (the hand head has n = 21 tokens, one per joint).
// vision_transformer.py:61-77:  S = Q K^T * 64^-0.5,  P = softmax(S),  O = P V   and its backward.
//
// A (b, h) problem is 21x21x64: far below a tensor-core tile and latency bound, so the design goal is threads in
// flight, not FLOPs.  A CTA of 4 warps owns one problem and the head dimension (64) is split across the warps:
// warp w works on d in [16w, 16w+16).  In the score phases lane j keeps its 16-wide slice of row j of K (or V) in
// registers and streams Q (or dO) rows from shared memory as broadcast float4 loads; the four partial score
// matrices are combined in shared memory; softmax reductions are warp shuffles; in the output phases lane i
// accumulates a 16-wide slice of row i.  Registers stay ~64/thread, so 8-12 CTAs are resident per SM.
#include "kernels.h"

namespace scat {
namespace {

constexpr int DH = 64;
constexpr int NW = 4;            // warps per CTA, each owns DS = 16 of the 64 head dims
constexpr int DS = DH / NW;

// [N][64] global rows -> registers (load phase) -> shared (store phase); compile-time trip count so that every
// global load of every tile is in flight before the first is consumed
template <int N>
struct RowRegs { float4 v[(N * (DH / 4) + NW * 32 - 1) / (NW * 32)]; };
template <int N>
__device__ __forceinline__ void rows_load(const float* __restrict__ src, long long row_stride, RowRegs<N>& t) {
#pragma unroll
    for (int i = 0; i < (N * (DH / 4) + NW * 32 - 1) / (NW * 32); ++i) {
        const int e = threadIdx.x + i * NW * 32, r = e >> 4, c4 = e & 15;
        t.v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < N) t.v[i] = __ldg(reinterpret_cast<const float4*>(src + (long long)r * row_stride) + c4);
    }
}
template <int N>
__device__ __forceinline__ void rows_store(float* dst, const RowRegs<N>& t) {
#pragma unroll
    for (int i = 0; i < (N * (DH / 4) + NW * 32 - 1) / (NW * 32); ++i) {
        const int e = threadIdx.x + i * NW * 32;
        if (e < N * (DH / 4)) reinterpret_cast<float4*>(dst)[e] = t.v[i];
    }
}
__device__ __forceinline__ void load_slice(const float* __restrict__ src, float* reg) {
#pragma unroll
    for (int c4 = 0; c4 < DS / 4; ++c4) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(src) + c4);
        reg[4 * c4] = v.x; reg[4 * c4 + 1] = v.y; reg[4 * c4 + 2] = v.z; reg[4 * c4 + 3] = v.w;
    }
}
__device__ __forceinline__ float dot_slice(const float* srow, const float* reg) {
    float a = 0.f;
#pragma unroll
    for (int c4 = 0; c4 < DS / 4; ++c4) {
        const float4 q = reinterpret_cast<const float4*>(srow)[c4];
        a = fmaf(q.x, reg[4 * c4], a); a = fmaf(q.y, reg[4 * c4 + 1], a);
        a = fmaf(q.z, reg[4 * c4 + 2], a); a = fmaf(q.w, reg[4 * c4 + 3], a);
    }
    return a;
}
__device__ __forceinline__ void axpy_slice(float w, const float* srow, float* acc) {
#pragma unroll
    for (int c4 = 0; c4 < DS / 4; ++c4) {
        const float4 v = reinterpret_cast<const float4*>(srow)[c4];
        acc[4 * c4] = fmaf(w, v.x, acc[4 * c4]); acc[4 * c4 + 1] = fmaf(w, v.y, acc[4 * c4 + 1]);
        acc[4 * c4 + 2] = fmaf(w, v.z, acc[4 * c4 + 2]); acc[4 * c4 + 3] = fmaf(w, v.w, acc[4 * c4 + 3]);
    }
}
// DS consecutive outputs starting at element `off` of a tensor stored in `out_mode` (fp32 / TF32-rounded / bf16)
__device__ __forceinline__ void store_slice(float* base, long long off, const float* acc, int out_mode) {
#pragma unroll
    for (int c4 = 0; c4 < DS / 4; ++c4)
        store_out4(base, off + 4 * c4, make_float4(acc[4 * c4], acc[4 * c4 + 1], acc[4 * c4 + 2], acc[4 * c4 + 3]), out_mode);
}

template <int N>
__global__ void __launch_bounds__(NW * 32)
attention_fwd_small_kernel(const float* __restrict__ QKV, float* __restrict__ O, float* __restrict__ P, int heads,
                           int round_out) {
    pdl_sync();
    constexpr int LS = N + 1;
    __shared__ __align__(16) float Qs[N * DH];
    __shared__ __align__(16) float Vs[N * DH];
    __shared__ float Sp[NW][N * LS];
    __shared__ float Ps[N * LS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int prob = blockIdx.x, b = prob / heads, g = prob % heads;
    const int inner = heads * DH;
    const long long rs = 3LL * inner;
    const float* base = QKV + (long long)b * N * rs + g * DH;
    const bool act = lane < N;
    float kreg[DS];
    {
        RowRegs<N> tq, tv;
        rows_load<N>(base, rs, tq);
        rows_load<N>(base + 2 * inner, rs, tv);
        if (act) load_slice(base + inner + (long long)lane * rs + warp * DS, kreg);
        rows_store<N>(Qs, tq);
        rows_store<N>(Vs, tv);
    }
    __syncthreads();
    if (act) {
#pragma unroll
        for (int i = 0; i < N; ++i) Sp[warp][i * LS + lane] = dot_slice(Qs + i * DH + warp * DS, kreg);
    }
    __syncthreads();
    float* Pg = P + (long long)prob * N * N;
    for (int i = warp; i < N; i += NW) {
        const float s = act ? (Sp[0][i * LS + lane] + Sp[1][i * LS + lane] + Sp[2][i * LS + lane] + Sp[3][i * LS + lane]) * 0.125f
                            : -INFINITY;
        const float m = warp_max(s);
        const float e = act ? expf(s - m) : 0.f;
        const float sum = warp_sum(e);
        const float p = e * (1.0f / sum);
        if (act) {
            Ps[i * LS + lane] = p;
            Pg[i * N + lane] = p;
        }
    }
    __syncthreads();
    if (act) {
        float acc[DS];
#pragma unroll
        for (int d = 0; d < DS; ++d) acc[d] = 0.f;
#pragma unroll 7
        for (int j = 0; j < N; ++j) axpy_slice(Ps[lane * LS + j], Vs + j * DH + warp * DS, acc);
        store_slice(O, ((long long)b * N + lane) * inner + g * DH + warp * DS, acc, round_out);
    }
}

template <int N>
__global__ void __launch_bounds__(NW * 32)
attention_bwd_small_kernel(const float* __restrict__ QKV, const float* __restrict__ P, const float* __restrict__ dO,
                           float* __restrict__ dQKV, int heads, int round_out, int act_batch) {
    pdl_sync();
    constexpr int LS = N + 1;
    __shared__ __align__(16) float Qs[N * DH];
    __shared__ __align__(16) float Ks[N * DH];
    __shared__ __align__(16) float Gs[N * DH];   // dO
    __shared__ float Sp[NW][N * LS];             // partial dP per d-slice
    __shared__ float Ps[N * LS];                 // P
    __shared__ float Ds[N * LS];                 // dS (scale folded in)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int prob = blockIdx.x, b = prob / heads, g = prob % heads;
    const int inner = heads * DH;
    const long long rs = 3LL * inner;
    const int ba = act_batch > 0 ? b % act_batch : b;        // stacked cotangents share the saved activations
    const float* base = QKV + (long long)ba * N * rs + g * DH;
    const float* Pg = P + ((long long)ba * heads + g) * N * N;
    const bool act = lane < N;
    float vreg[DS];
    {
        RowRegs<N> tq, tk, tg;
        float pr[(N * N + NW * 32 - 1) / (NW * 32)];
        rows_load<N>(base, rs, tq);
        rows_load<N>(base + inner, rs, tk);
        rows_load<N>(dO + (long long)b * N * inner + g * DH, inner, tg);
#pragma unroll
        for (int i = 0; i < (N * N + NW * 32 - 1) / (NW * 32); ++i) {
            const int e = threadIdx.x + i * NW * 32;
            pr[i] = e < N * N ? __ldg(Pg + e) : 0.f;
        }
        if (act) load_slice(base + 2 * inner + (long long)lane * rs + warp * DS, vreg);
        rows_store<N>(Qs, tq);
        rows_store<N>(Ks, tk);
        rows_store<N>(Gs, tg);
#pragma unroll
        for (int i = 0; i < (N * N + NW * 32 - 1) / (NW * 32); ++i) {
            const int e = threadIdx.x + i * NW * 32;
            if (e < N * N) Ps[(e / N) * LS + (e % N)] = pr[i];
        }
    }
    __syncthreads();
    if (act) {
#pragma unroll
        for (int i = 0; i < N; ++i) Sp[warp][i * LS + lane] = dot_slice(Gs + i * DH + warp * DS, vreg);   // dP[i,j] slice
    }
    __syncthreads();
    for (int i = warp; i < N; i += NW) {
        const float dp = act ? Sp[0][i * LS + lane] + Sp[1][i * LS + lane] + Sp[2][i * LS + lane] + Sp[3][i * LS + lane] : 0.f;
        const float p = act ? Ps[i * LS + lane] : 0.f;
        const float r = warp_sum(dp * p);                        // rowsum(dP * P)
        if (act) Ds[i * LS + lane] = p * (dp - r) * 0.125f;      // dS
    }
    __syncthreads();
    if (act) {
        const long long drow = ((long long)b * N + lane) * rs + g * DH + warp * DS;
        float acc[DS];
#pragma unroll
        for (int d = 0; d < DS; ++d) acc[d] = 0.f;
#pragma unroll 7
        for (int i = 0; i < N; ++i) axpy_slice(Ps[i * LS + lane], Gs + i * DH + warp * DS, acc);    // dV[j] = sum_i P[i,j] dO[i]
        store_slice(dQKV, drow + 2 * inner, acc, round_out);
#pragma unroll
        for (int d = 0; d < DS; ++d) acc[d] = 0.f;
#pragma unroll 7
        for (int i = 0; i < N; ++i) axpy_slice(Ds[i * LS + lane], Qs + i * DH + warp * DS, acc);    // dK[j] = sum_i dS[i,j] Q[i]
        store_slice(dQKV, drow + inner, acc, round_out);
#pragma unroll
        for (int d = 0; d < DS; ++d) acc[d] = 0.f;
#pragma unroll 7
        for (int j = 0; j < N; ++j) axpy_slice(Ds[lane * LS + j], Ks + j * DH + warp * DS, acc);    // dQ[i] = sum_j dS[i,j] K[j]
        store_slice(dQKV, drow, acc, round_out);
    }
}

}  // namespace

// n == 21 (one token per joint, hand_net.py:328) is the only short length on the reference path
bool attention_small_supported(int n) { return n == 21; }
int launch_attention_small_fwd(const float* QKV, float* O, float* P, int B, int n, int heads, int round_out,
                               cudaStream_t stream) {
    SCAT_REQUIRE(n == 21, kErrUnsupported, "attention_small: n=%d", n);
    SCAT_CHECK_CUDA(launch_k(attention_fwd_small_kernel<21>, dim3(B * heads), dim3(NW * 32), 0, stream, QKV, O, P, heads, round_out));
    SCAT_CHECK_LAUNCH();
    return 0;
}
int launch_attention_small_bwd(const float* QKV, const float* P, const float* dO, float* dQKV, int B, int n, int heads,
                               int round_out, cudaStream_t stream, int act_batch) {
    SCAT_REQUIRE(n == 21, kErrUnsupported, "attention_small: n=%d", n);
    SCAT_CHECK_CUDA(launch_k(attention_bwd_small_kernel<21>, dim3(B * heads), dim3(NW * 32), 0, stream, QKV, P, dO, dQKV, heads, round_out, act_batch));
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace scat
