// Front end of the head: 1x1 channel-reduction conv + positional encoding + token masking, and its
// backward (hand_net.py:329,363-373).  HBM-bound: the conv reads x2 (1.6 MB/sample), its dgrad writes
// x2.grad (1.6 MB/sample) and its wgrad reads x2 again.
//
//   Fv[b,t,p] = sum_c Wc[t,c] x2[b,c,p]              feat_visual (returned to the caller)
//   X0[b,t,p] = Fv + pe[t,p]        (pos_embed)      token matrix fed to the transformer
//   X0[b,idx,:] = mask_token                         idx: host-drawn indices, same for every sample
//
// fp32 CUDA-core version: each thread owns 4 consecutive pixels x all T (<=24) tokens, warps split the
// channel reduction, partial sums are combined through shared memory.
#include "kernels.h"

namespace scat {
namespace {

constexpr int TP = 24;          // token count padded to a float4 multiple in shared memory
constexpr int CONV_THREADS = 256;
constexpr int CONV_WARPS = CONV_THREADS / 32;

__device__ __forceinline__ uint32_t build_mask_bits(const int32_t* __restrict__ mask_idx, int n_masked,
                                                    uint32_t* s_bits /* [4] zeroed */) {
    // token masks for up to 128 tokens in 4 words of shared memory; returns nothing useful, call then sync
    for (int i = threadIdx.x; i < n_masked; i += blockDim.x) {
        const int t = mask_idx[i];
        if (t >= 0 && t < 128) atomicOr(&s_bits[t >> 5], 1u << (t & 31));
    }
    return 0;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int T>
__global__ void __launch_bounds__(CONV_THREADS, 2)
conv_pe_mask_fwd_kernel(const float* __restrict__ x2, const float* __restrict__ Wc, const float* __restrict__ pe,
                        const float* __restrict__ mask_token, const int32_t* __restrict__ mask_idx, int n_masked,
                        int pos_embed, float* __restrict__ Fv, float* __restrict__ X0, int B, int C, int HW) {
    pdl_sync();
    extern __shared__ __align__(16) float smem[];
    __shared__ uint32_t s_bits[4];
    float* Ws = smem;  // [C][TP], later reused as the cross-warp reduction buffer [warps][T][4][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 4) s_bits[tid] = 0;
    for (int i = tid; i < C * TP; i += CONV_THREADS) {
        const int c = i / TP, t = i % TP;
        Ws[i] = t < T ? Wc[t * C + c] : 0.f;
    }
    __syncthreads();
    build_mask_bits(mask_idx, n_masked, s_bits);

    const int groups_per_sample = HW >> 2;
    const long long total_groups = (long long)B * groups_per_sample;
    const long long grp = (long long)blockIdx.x * 32 + lane;
    const bool active = grp < total_groups;
    const int b = active ? (int)(grp / groups_per_sample) : 0;
    const int p = active ? (int)(grp % groups_per_sample) * 4 : 0;

    float acc[T][4];
#pragma unroll
    for (int t = 0; t < T; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;

    const int c_per_warp = (C + CONV_WARPS - 1) / CONV_WARPS;
    const int c_beg = warp * c_per_warp, c_end = min(C, c_beg + c_per_warp);
    const float* xb = x2 + ((long long)b * C) * HW + p;
    constexpr int U = 4;
    for (int c0 = c_beg; c0 < c_end; c0 += U) {
        float4 xv[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            xv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (active && c0 + u < c_end) xv[u] = __ldg(reinterpret_cast<const float4*>(xb + (long long)(c0 + u) * HW));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (c0 + u < c_end) {
                const float4* wrow = reinterpret_cast<const float4*>(Ws + (c0 + u) * TP);
#pragma unroll
                for (int t4 = 0; t4 < (T + 3) / 4; ++t4) {
                    const float4 w = wrow[t4];
                    const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const int t = t4 * 4 + k;
                        if (t < T) {
                            acc[t][0] = fmaf(wv[k], xv[u].x, acc[t][0]);
                            acc[t][1] = fmaf(wv[k], xv[u].y, acc[t][1]);
                            acc[t][2] = fmaf(wv[k], xv[u].z, acc[t][2]);
                            acc[t][3] = fmaf(wv[k], xv[u].w, acc[t][3]);
                        }
                    }
                }
            }
        }
    }
    __syncthreads();  // everyone is done reading Ws; reuse it for the reduction
    float* red = smem;  // [warp][T][4][32]
#pragma unroll
    for (int t = 0; t < T; ++t)
#pragma unroll
        for (int q = 0; q < 4; ++q) red[((warp * T + t) * 4 + q) * 32 + lane] = acc[t][q];
    __syncthreads();

    for (int o = tid; o < T * 32; o += CONV_THREADS) {
        const int t = o >> 5, l = o & 31;
        const long long g2 = (long long)blockIdx.x * 32 + l;
        if (g2 >= total_groups) continue;
        float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int w = 0; w < CONV_WARPS; ++w)
#pragma unroll
            for (int q = 0; q < 4; ++q) s[q] += red[((w * T + t) * 4 + q) * 32 + l];
        const int bb = (int)(g2 / groups_per_sample);
        const int pp = (int)(g2 % groups_per_sample) * 4;
        const long long off = ((long long)bb * T + t) * HW + pp;
        const float4 fv = make_float4(s[0], s[1], s[2], s[3]);
        *reinterpret_cast<float4*>(Fv + off) = fv;
        float4 x0 = fv;
        if ((s_bits[t >> 5] >> (t & 31)) & 1u) {
            x0 = __ldg(reinterpret_cast<const float4*>(mask_token + pp));
        } else if (pos_embed) {
            const float4 e = __ldg(reinterpret_cast<const float4*>(pe + (long long)t * HW + pp));
            x0 = make_float4(fv.x + e.x, fv.y + e.y, fv.z + e.z, fv.w + e.w);
        }
        if (X0 != Fv) *reinterpret_cast<float4*>(X0 + off) = x0;
        else if ((s_bits[t >> 5] >> (t & 31)) & 1u) *reinterpret_cast<float4*>(Fv + off) = x0;  // aliased overwrite
    }
}

// X0 = tokens (+ pe) with masked rows replaced; one float per thread (config 4 front end)
__global__ void pe_mask_tokens_kernel(const float* __restrict__ tok, const float* __restrict__ pe,
                                      const float* __restrict__ mask_token, const int32_t* __restrict__ mask_idx,
                                      int n_masked, int pos_embed, float* __restrict__ X0, int B, int T, int D) {
    pdl_sync();
    __shared__ uint32_t s_bits[4];
    if (threadIdx.x < 4) s_bits[threadIdx.x] = 0;
    __syncthreads();
    build_mask_bits(mask_idx, n_masked, s_bits);
    __syncthreads();
    const long long total = (long long)B * T * D;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        const int t = (int)((i / D) % T);
        float v;
        if ((s_bits[t >> 5] >> (t & 31)) & 1u) v = mask_token[d];
        else v = tok[i] + (pos_embed ? pe[(long long)t * D + d] : 0.f);
        X0[i] = v;
    }
}

// dFv = dX0 with masked rows zeroed (float4 per thread)
__global__ void mask_bwd_copy_kernel(const float* __restrict__ dX0, const int32_t* __restrict__ mask_idx, int n_masked,
                                     int keep_masked, float* __restrict__ dFv, int B, int T, int HW) {
    pdl_sync();
    __shared__ uint32_t s_bits[4];
    if (threadIdx.x < 4) s_bits[threadIdx.x] = 0;
    __syncthreads();
    build_mask_bits(mask_idx, n_masked, s_bits);
    __syncthreads();
    const int hw4 = HW >> 2;
    const long long total = (long long)B * T * hw4;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)((i / hw4) % T);
        float4 v = reinterpret_cast<const float4*>(dX0)[i];
        if (!keep_masked && ((s_bits[t >> 5] >> (t & 31)) & 1u)) v = make_float4(0.f, 0.f, 0.f, 0.f);
        reinterpret_cast<float4*>(dFv)[i] = v;
    }
}

// d mask_token[p] = sum_b sum_{t in idx} dX0[b,t,p]: grid (pixel blocks, sample slices), slices combine with atomics
__global__ void mask_token_grad_kernel(const float* __restrict__ dX0, const int32_t* __restrict__ mask_idx,
                                       int n_masked, float* __restrict__ d_mask_token, int B, int T, int HW) {
    pdl_sync();
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= HW) return;
    float s = 0.f;
    for (int b = blockIdx.y; b < B; b += gridDim.y)
        for (int i = 0; i < n_masked; ++i) s += dX0[((long long)b * T + mask_idx[i]) * HW + p];
    atomicAdd(d_mask_token + p, s);
}

// x2g[b,c,p] = sum_t Wc[t,c] dFv[b,t,p]
template <int T>
__global__ void __launch_bounds__(CONV_THREADS, 2)
conv_dgrad_kernel(const float* __restrict__ dFv, const float* __restrict__ Wc, float* __restrict__ x2g, int B, int C,
                  int HW) {
    pdl_sync();
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;  // [C][TP]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < C * TP; i += CONV_THREADS) {
        const int c = i / TP, t = i % TP;
        Ws[i] = t < T ? Wc[t * C + c] : 0.f;
    }
    __syncthreads();
    const int groups_per_sample = HW >> 2;
    const long long total_groups = (long long)B * groups_per_sample;
    const long long grp = (long long)blockIdx.x * 32 + lane;
    if (grp >= total_groups) return;
    const int b = (int)(grp / groups_per_sample);
    const int p = (int)(grp % groups_per_sample) * 4;
    float4 d[T];
#pragma unroll
    for (int t = 0; t < T; ++t) d[t] = __ldg(reinterpret_cast<const float4*>(dFv + ((long long)b * T + t) * HW + p));
    const int c_per_warp = (C + CONV_WARPS - 1) / CONV_WARPS;
    const int c_beg = warp * c_per_warp, c_end = min(C, c_beg + c_per_warp);
    float* ob = x2g + ((long long)b * C) * HW + p;
#pragma unroll 2
    for (int c = c_beg; c < c_end; ++c) {
        const float4* wrow = reinterpret_cast<const float4*>(Ws + c * TP);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int t4 = 0; t4 < (T + 3) / 4; ++t4) {
            const float4 w = wrow[t4];
            const float wv[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int t = t4 * 4 + k;
                if (t < T) {
                    o.x = fmaf(wv[k], d[t].x, o.x);
                    o.y = fmaf(wv[k], d[t].y, o.y);
                    o.z = fmaf(wv[k], d[t].z, o.z);
                    o.w = fmaf(wv[k], d[t].w, o.w);
                }
            }
        }
        __stcs(reinterpret_cast<float4*>(ob + (long long)c * HW), o);   // streaming store: never re-read here
    }
}

// dWc[t,c] = sum_{b,p} dFv[b,t,p] x2[b,c,p]: persistent CTAs accumulate a private [T][C] partial over
// their share of (sample, pixel-chunk) work items, then a second kernel sums the partials (deterministic).
constexpr int kConvWgradCtas = 148 * 2;
constexpr int WG_PX = 28;       // pixels per work item (one image row of the 28x28 map)
template <int T, int CPT /* channels per thread */>
__global__ void __launch_bounds__(CONV_THREADS, 2)
conv_wgrad_partial_kernel(const float* __restrict__ dFv, const float* __restrict__ x2, float* __restrict__ partial,
                          int B, int C, int HW) {
    pdl_sync();
    extern __shared__ __align__(16) float smem[];
    float* xs = smem;                  // [C][WG_PX]
    float* ds = smem + C * WG_PX;      // [T][WG_PX]
    const int tid = threadIdx.x;
    const int chunks = HW / WG_PX;
    const int items = B * chunks;
    float acc[CPT][T];
#pragma unroll
    for (int i = 0; i < CPT; ++i)
#pragma unroll
        for (int t = 0; t < T; ++t) acc[i][t] = 0.f;
    const int px4 = WG_PX / 4;
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        const int b = item / chunks, p0 = (item % chunks) * WG_PX;
        __syncthreads();
        // stage the [C][28] slab of x2 and the [T][28] slab of dFv with cp.async: every 16-byte copy is in flight
        // at once and none of them holds a register (a load->store loop would serialise one round trip per pass)
        for (int i = tid; i < C * px4; i += CONV_THREADS) {
            const int c = i / px4, q = i % px4;
            cp_async16(reinterpret_cast<float4*>(xs) + i, reinterpret_cast<const float4*>(x2 + ((long long)b * C + c) * HW + p0) + q);
        }
        for (int i = tid; i < T * px4; i += CONV_THREADS) {
            const int t = i / px4, q = i % px4;
            cp_async16(reinterpret_cast<float4*>(ds) + i, reinterpret_cast<const float4*>(dFv + ((long long)b * T + t) * HW + p0) + q);
        }
        cp_async_wait_all();
        __syncthreads();
#pragma unroll 1
        for (int q = 0; q < px4; ++q) {
            float4 xv[CPT];
#pragma unroll
            for (int i = 0; i < CPT; ++i) {
                const int c = tid + i * CONV_THREADS;
                xv[i] = c < C ? reinterpret_cast<const float4*>(xs + c * WG_PX)[q] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int t = 0; t < T; ++t) {
                const float4 dv = reinterpret_cast<const float4*>(ds + t * WG_PX)[q];
#pragma unroll
                for (int i = 0; i < CPT; ++i) {
                    acc[i][t] = fmaf(dv.x, xv[i].x, acc[i][t]);
                    acc[i][t] = fmaf(dv.y, xv[i].y, acc[i][t]);
                    acc[i][t] = fmaf(dv.z, xv[i].z, acc[i][t]);
                    acc[i][t] = fmaf(dv.w, xv[i].w, acc[i][t]);
                }
            }
        }
    }
    float* out = partial + (long long)blockIdx.x * T * C;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
        const int c = tid + i * CONV_THREADS;
        if (c < C)
#pragma unroll
            for (int t = 0; t < T; ++t) out[t * C + c] = acc[i][t];
    }
}

__global__ void reduce_partials_kernel(const float* __restrict__ partial, int n_parts, int n, float* __restrict__ out) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int k = 0; k < n_parts; ++k) s += partial[(long long)k * n + i];
    out[i] = s;
}

}  // namespace

int launch_conv_pe_mask_fwd(const float* x2, const float* Wc, const float* pe, const float* mask_token,
                            const int32_t* mask_idx, int n_masked, int pos_embed, float* feat_visual, float* X0,
                            int B, int C, int HW, int T, cudaStream_t stream) {
    SCAT_REQUIRE(T == 21, kErrUnsupported, "conv fwd: only T=21 tokens is built (got %d)", T);
    SCAT_REQUIRE(HW % 4 == 0 && C % CONV_WARPS == 0, kErrUnsupported, "conv fwd: HW%%4 / C%%8 (HW=%d C=%d)", HW, C);
    SCAT_REQUIRE(n_masked == 0 || mask_idx != nullptr, kErrBadArg, "conv fwd: mask_idx is null");
    const size_t smem = sizeof(float) * (size_t)max(C * TP, CONV_WARPS * T * 4 * 32);
    SCAT_ENSURE_SMEM(conv_pe_mask_fwd_kernel<21>, 96 * 1024);
    SCAT_REQUIRE(smem <= 96 * 1024, kErrUnsupported, "conv fwd: C too large");
    const long long groups = (long long)B * (HW / 4);
    const int grid = (int)((groups + 31) / 32);
    SCAT_CHECK_CUDA(launch_k(conv_pe_mask_fwd_kernel<21>, dim3(grid), dim3(CONV_THREADS), smem, stream, x2, Wc, pe, mask_token, mask_idx, n_masked,
                                                                      pos_embed, feat_visual, X0, B, C, HW));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_pe_mask_tokens(const float* tokens, const float* pe, const float* mask_token, const int32_t* mask_idx,
                          int n_masked, int pos_embed, float* X0, int B, int T, int D, cudaStream_t stream) {
    SCAT_REQUIRE(T <= 128, kErrUnsupported, "token front end: at most 128 tokens (got %d)", T);
    const long long total = (long long)B * T * D;
    const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    SCAT_CHECK_CUDA(launch_k(pe_mask_tokens_kernel, dim3(grid), dim3(256), 0, stream, tokens, pe, mask_token, mask_idx, n_masked, pos_embed, X0, B, T, D));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_mask_bwd(const float* dX0, const int32_t* mask_idx, int n_masked, int keep_masked, float* dFv,
                    float* d_mask_token, int B, int T, int HW, cudaStream_t stream, int token_grad_zeroed) {
    SCAT_REQUIRE(HW % 4 == 0 && T <= 128, kErrUnsupported, "mask bwd: HW%%4, T<=128");
    if (dFv != nullptr) {
        const long long total = (long long)B * T * (HW / 4);
        const int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
        SCAT_CHECK_CUDA(launch_k(mask_bwd_copy_kernel, dim3(grid), dim3(256), 0, stream, dX0, mask_idx, n_masked, keep_masked, dFv, B, T, HW));
        SCAT_CHECK_LAUNCH();
    }
    if (d_mask_token != nullptr) {
        if (!token_grad_zeroed) SCAT_CHECK_CUDA(cudaMemsetAsync(d_mask_token, 0, (size_t)HW * sizeof(float), stream));
        if (n_masked > 0) {
            SCAT_CHECK_CUDA(launch_k(mask_token_grad_kernel, dim3(dim3(ceil_div(HW, 128), min(B, 32))), dim3(128), 0, stream, dX0, mask_idx, n_masked,
                                                                                            d_mask_token, B, T, HW));
            SCAT_CHECK_LAUNCH();
        }
    }
    return 0;
}

int launch_conv_dgrad(const float* dFv, const float* Wc, float* x2_grad, int B, int C, int HW, int T,
                      cudaStream_t stream) {
    SCAT_REQUIRE(T == 21 && HW % 4 == 0, kErrUnsupported, "conv dgrad: T=21, HW%%4");
    const size_t smem = sizeof(float) * (size_t)C * TP;
    SCAT_ENSURE_SMEM(conv_dgrad_kernel<21>, 96 * 1024);
    SCAT_REQUIRE(smem <= 96 * 1024, kErrUnsupported, "conv dgrad: C too large");
    const long long groups = (long long)B * (HW / 4);
    SCAT_CHECK_CUDA(launch_k(conv_dgrad_kernel<21>, dim3((int)((groups + 31) / 32)), dim3(CONV_THREADS), smem, stream, dFv, Wc, x2_grad, B, C, HW));
    SCAT_CHECK_LAUNCH();
    return 0;
}


size_t conv_wgrad_scratch_floats(int C, int T) { return (size_t)kConvWgradCtas * T * C; }

int launch_conv_wgrad(const float* dFv, const float* x2, float* dWc, float* scratch, int B, int C, int HW, int T,
                      cudaStream_t stream) {
    SCAT_REQUIRE(T == 21 && HW % WG_PX == 0 && C <= 2 * CONV_THREADS, kErrUnsupported,
                 "conv wgrad: T=21, HW%%28, C<=512 (T=%d HW=%d C=%d)", T, HW, C);
    SCAT_REQUIRE(scratch != nullptr, kErrBadArg, "conv wgrad: scratch is null");
    const size_t smem = sizeof(float) * ((size_t)C * WG_PX + (size_t)T * WG_PX);
    SCAT_ENSURE_SMEM((conv_wgrad_partial_kernel<21, 2>), 96 * 1024);
    const int items = B * (HW / WG_PX);
    const int grid = min(items, kConvWgradCtas);
    SCAT_CHECK_CUDA(launch_k(conv_wgrad_partial_kernel<21, 2>, dim3(grid), dim3(CONV_THREADS), smem, stream, dFv, x2, scratch, B, C, HW));
    SCAT_CHECK_LAUNCH();
    SCAT_CHECK_CUDA(launch_k(reduce_partials_kernel, dim3(ceil_div(T * C, 256)), dim3(256), 0, stream, scratch, grid, T * C, dWc));
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace scat
