// Softmax attention over a short token sequence (vision_transformer.py:61-77, mask=None):
//   per (sample b, head g):  S = (Q K^T) * 64^-0.5 ; P = softmax(S) ; O = P V
// QKV is the [B*n, 3*inner] output of to_qkv (q | k | v; head g = columns g*64 .. g*64+63, the
// 'b n (h d) -> b h n d' rearrange at :62), O is written head-merged [B*n, inner] (:77).
// n = 21 for the hand head (one token per joint), up to 128 for the HRNet-token variant.  At these
// sizes a (b,h) problem is far below one UMMA tile, so each CTA keeps Q,K,V in shared memory and
// uses FFMA; the kernel is latency/launch bound, not tensor bound.
#include <stdlib.h>

#include "kernels.h"

namespace scat {
namespace {

constexpr int DH = 64;        // dim_head (hand_net.py:331)
constexpr int LDS = DH + 1;   // padded row stride: conflict-free when lanes walk rows
constexpr int ATT_THREADS = 128;

__device__ __forceinline__ void load_head_tile(const float* __restrict__ src, long long row_stride, int n, float* dst) {
    // [n][64] global rows -> [n][65] shared
    for (int i = threadIdx.x; i < n * (DH / 4); i += blockDim.x) {
        const int r = i / (DH / 4), c4 = i % (DH / 4);
        const float4 v = __ldg(reinterpret_cast<const float4*>(src + (long long)r * row_stride) + c4);
        float* d = dst + r * LDS + c4 * 4;
        d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w;
    }
}

__global__ void __launch_bounds__(ATT_THREADS)
attention_fwd_kernel(const float* __restrict__ QKV, float* __restrict__ O, float* __restrict__ P, int n, int heads,
                     int round_out) {
    pdl_sync();
    extern __shared__ float sm[];
    float* Qs = sm;
    float* Ks = Qs + n * LDS;
    float* Vs = Ks + n * LDS;
    float* Ss = Vs + n * LDS;   // [n][n+1]
    const int lds = n + 1;
    const int b = blockIdx.x / heads, g = blockIdx.x % heads;
    const int inner = heads * DH;
    const long long rs = 3LL * inner;
    const float* base = QKV + (long long)b * n * rs + g * DH;
    load_head_tile(base, rs, n, Qs);
    load_head_tile(base + inner, rs, n, Ks);
    load_head_tile(base + 2 * inner, rs, n, Vs);
    __syncthreads();
    const float scale = 0.125f;   // 64^-0.5, applied after the dot product like the reference (:64)
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e % n;
        float s = 0.f;
#pragma unroll 16
        for (int d = 0; d < DH; ++d) s = fmaf(Qs[i * LDS + d], Ks[j * LDS + d], s);
        Ss[i * lds + j] = s * scale;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    float* Pg = P + (long long)blockIdx.x * n * n;
    for (int i = warp; i < n; i += nw) {
        float m = -INFINITY;
        for (int j = lane; j < n; j += 32) m = fmaxf(m, Ss[i * lds + j]);
        m = warp_max(m);
        float sum = 0.f;
        for (int j = lane; j < n; j += 32) {
            const float e = expf(Ss[i * lds + j] - m);
            Ss[i * lds + j] = e;
            sum += e;
        }
        sum = warp_sum(sum);
        const float inv = 1.0f / sum;
        for (int j = lane; j < n; j += 32) {
            const float p = Ss[i * lds + j] * inv;
            Ss[i * lds + j] = p;
            Pg[i * n + j] = p;
        }
    }
    __syncthreads();
    const long long ob = (long long)b * n * inner + g * DH;
    for (int e = threadIdx.x; e < n * DH; e += blockDim.x) {
        const int i = e / DH, d = e % DH;
        float o = 0.f;
        for (int j = 0; j < n; ++j) o = fmaf(Ss[i * lds + j], Vs[j * LDS + d], o);
        store_out(O, ob + (long long)i * inner + d, o, round_out);
    }
}

__global__ void __launch_bounds__(ATT_THREADS)
attention_bwd_kernel(const float* __restrict__ QKV, const float* __restrict__ P, const float* __restrict__ dO,
                     float* __restrict__ dQKV, int n, int heads, int round_out) {
    pdl_sync();
    extern __shared__ float sm[];
    float* Qs = sm;
    float* Ks = Qs + n * LDS;
    float* Vs = Ks + n * LDS;
    float* Gs = Vs + n * LDS;   // dO
    float* Ps = Gs + n * LDS;   // [n][n+1]
    float* Ds = Ps + n * (n + 1);  // dP then dS, [n][n+1]
    const int lds = n + 1;
    const int b = blockIdx.x / heads, g = blockIdx.x % heads;
    const int inner = heads * DH;
    const long long rs = 3LL * inner;
    const float* base = QKV + (long long)b * n * rs + g * DH;
    load_head_tile(base, rs, n, Qs);
    load_head_tile(base + inner, rs, n, Ks);
    load_head_tile(base + 2 * inner, rs, n, Vs);
    load_head_tile(dO + (long long)b * n * inner + g * DH, inner, n, Gs);
    const float* Pg = P + (long long)blockIdx.x * n * n;
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) Ps[(e / n) * lds + (e % n)] = Pg[e];
    __syncthreads();
    // dP = dO V^T
    for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
        const int i = e / n, j = e % n;
        float s = 0.f;
#pragma unroll 16
        for (int d = 0; d < DH; ++d) s = fmaf(Gs[i * LDS + d], Vs[j * LDS + d], s);
        Ds[i * lds + j] = s;
    }
    __syncthreads();
    // dS = P * (dP - rowsum(dP*P)) * scale
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int i = warp; i < n; i += nw) {
        float r = 0.f;
        for (int j = lane; j < n; j += 32) r = fmaf(Ds[i * lds + j], Ps[i * lds + j], r);
        r = warp_sum(r);
        for (int j = lane; j < n; j += 32) Ds[i * lds + j] = Ps[i * lds + j] * (Ds[i * lds + j] - r) * 0.125f;
    }
    __syncthreads();
    const long long dbase = (long long)b * n * rs + g * DH;
    for (int e = threadIdx.x; e < n * DH; e += blockDim.x) {
        const int i = e / DH, d = e % DH;
        float dq = 0.f, dk = 0.f, dv = 0.f;
        for (int j = 0; j < n; ++j) {
            dq = fmaf(Ds[i * lds + j], Ks[j * LDS + d], dq);   // dQ[i] = sum_j dS[i,j] K[j]
            dk = fmaf(Ds[j * lds + i], Qs[j * LDS + d], dk);   // dK[i] = sum_j dS[j,i] Q[j]
            dv = fmaf(Ps[j * lds + i], Gs[j * LDS + d], dv);   // dV[i] = sum_j P[j,i] dO[j]
        }
        const long long o = dbase + (long long)i * rs + d;
        store_out(dQKV, o, dq, round_out);
        store_out(dQKV, o + inner, dk, round_out);
        store_out(dQKV, o + 2 * inner, dv, round_out);
    }
}

}  // namespace

// attention_small.cu: fp32 FFMA kernels for the n = 21 training path
bool attention_small_supported(int n);
int launch_attention_small_fwd(const float* QKV, float* O, float* P, int B, int n, int heads, int round_out,
                               cudaStream_t stream);
int launch_attention_small_bwd(const float* QKV, const float* P, const float* dO, float* dQKV, int B, int n, int heads,
                               int round_out, cudaStream_t stream, int act_batch);

// attention_mma.cu: mma.sync TF32 kernels for n = 21 (TF32 / BF16 precisions)
int launch_attention_mma_fwd(const float* QKV, float* O, float* P, int B, int n, int heads, int out_mode, cudaStream_t stream);
int launch_attention_mma_bwd(const float* QKV, const float* P, const float* dO, float* dQKV, int B, int n, int heads,
                             int out_mode, cudaStream_t stream, int act_batch);

// attention_tc128.cu: tcgen05 kernel for n = 128 (config 4, inference: P is not written)
bool attention_tc128_supported(int n, int heads, const float* QKV);
int launch_attention_tc128_fwd(const float* QKV, float* O, int B, int n, int heads, int out_mode, cudaStream_t stream);

int launch_attention_fwd(const float* QKV, float* O, float* P, int B, int n, int heads, int round_out,
                         cudaStream_t stream) {
    SCAT_REQUIRE(n >= 1 && n <= 128, kErrUnsupported, "attention: n=%d not in [1,128]", n);
    if (n == 21 && round_out != OUT_F32) return launch_attention_mma_fwd(QKV, O, P, B, n, heads, round_out, stream);
    // n = 128 only occurs on the inference-only token path (hand_net.py:193-203): no backward, so P is not needed
    if (round_out != OUT_F32 && attention_tc128_supported(n, heads, QKV)) return launch_attention_tc128_fwd(QKV, O, B, n, heads, round_out, stream);
    if (attention_small_supported(n)) return launch_attention_small_fwd(QKV, O, P, B, n, heads, round_out, stream);
    const size_t smem = sizeof(float) * ((size_t)3 * n * LDS + (size_t)n * (n + 1));
    if (smem > 48 * 1024) SCAT_ENSURE_SMEM(attention_fwd_kernel, smem);
    SCAT_CHECK_CUDA(launch_k(attention_fwd_kernel, dim3(B * heads), dim3(ATT_THREADS), smem, stream, QKV, O, P, n, heads, round_out));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_attention_bwd(const float* QKV, const float* P, const float* dO, float* dQKV, int B, int n, int heads,
                         int round_out, cudaStream_t stream, int act_batch) {
    SCAT_REQUIRE(n >= 1 && n <= 64, kErrUnsupported, "attention bwd: n=%d not in [1,64] (training path is n=21)", n);
    if (n == 21 && round_out != OUT_F32) return launch_attention_mma_bwd(QKV, P, dO, dQKV, B, n, heads, round_out, stream, act_batch);
    if (attention_small_supported(n)) return launch_attention_small_bwd(QKV, P, dO, dQKV, B, n, heads, round_out, stream, act_batch);
    SCAT_REQUIRE(act_batch == 0, kErrUnsupported, "attention bwd: stacked cotangents only on the n=21 path");
    const size_t smem = sizeof(float) * ((size_t)4 * n * LDS + (size_t)2 * n * (n + 1));
    if (smem > 48 * 1024) SCAT_ENSURE_SMEM(attention_bwd_kernel, smem);
    SCAT_CHECK_CUDA(launch_k(attention_bwd_kernel, dim3(B * heads), dim3(ATT_THREADS), smem, stream, QKV, P, dO, dQKV, n, heads, round_out));
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace scat
