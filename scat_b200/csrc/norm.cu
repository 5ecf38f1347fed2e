// LayerNorm forward / backward over token rows (vision_transformer.py:23,26: nn.LayerNorm(dim), eps=1e-5,
// affine).  One warp per row, the row lives in registers; two-pass mean/variance like ATen's CPU kernel.
#include "kernels.h"

namespace scat {
namespace {

constexpr float kEps = 1e-5f;
constexpr int LN_WARPS = 8;

template <int NPL>  // elements per lane (row length <= 32*NPL)
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_fwd_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ Y, int ldy, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int M, int D, int round_out) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* x = X + (long long)row * ldx;
    float v[NPL];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int c = i * 32 + lane;
        v[i] = c < D ? x[c] : 0.f;
        s += v[i];
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int c = i * 32 + lane;
        const float d = c < D ? v[i] - mean : 0.f;
        q = fmaf(d, d, q);
    }
    const float var = warp_sum(q) / (float)D;
    const float rstd = 1.0f / sqrtf(var + kEps);
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int c = i * 32 + lane;
        if (c < D) {
            const float o = (v[i] - mean) * rstd * gamma[c] + beta[c];
            store_out(Y, (long long)row * ldy + c, o, round_out);
        }
    }
    if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
}

template <int NPL>
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_bwd_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ X, int ldx,
                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ resid, int ldr, float* __restrict__ dX, int lddx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int D, int round_out, int act_rows,
                     __nv_bfloat16* __restrict__ dX16, int lddx16) {
    pdl_sync();
    __shared__ float red[LN_WARPS][32 * NPL];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float dg[NPL], db[NPL];
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        dg[i] = 0.f;
        db[i] = 0.f;
    }
    const float invD = 1.0f / (float)D;
    for (int row = blockIdx.x * LN_WARPS + warp; row < M; row += gridDim.x * LN_WARPS) {
        const int arow = act_rows > 0 ? row % act_rows : row;
        const bool real = act_rows <= 0 || row < act_rows;     // stacked ones-cotangent rows carry no parameter gradient
        const float mu = mean[arow], rs = rstd[arow];
        const float* x = X + (long long)arow * ldx;
        const float* dy = dY + (long long)row * lddy;
        const float* rr = resid != nullptr ? resid + (long long)row * ldr : nullptr;
        // every global read of the row (x, dy, residual) is issued before the first store: dX may alias any of them
        // as far as the compiler knows, and a load placed after a store waits for one L2 round trip per element
        float xh[NPL], g[NPL], rv[NPL];
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const int c = i * 32 + lane;
            xh[i] = c < D ? __ldg(x + c) : mu;
            g[i] = c < D ? __ldg(dy + c) : 0.f;
            rv[i] = (rr != nullptr && c < D) ? __ldg(rr + c) : 0.f;
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const int c = i * 32 + lane;
            const float dv = g[i];
            xh[i] = (xh[i] - mu) * rs;
            g[i] = dv * (c < D ? __ldg(gamma + c) : 0.f);
            s1 += g[i];
            s2 = fmaf(g[i], xh[i], s2);
            if (real) {
                dg[i] = fmaf(dv, xh[i], dg[i]);
                db[i] += dv;
            }
        }
        s1 = warp_sum(s1) * invD;
        s2 = warp_sum(s2) * invD;
        float* dx = dX + (long long)row * lddx;
#pragma unroll
        for (int i = 0; i < NPL; ++i) {
            const int c = i * 32 + lane;
            if (c < D) {
                const float o = rs * (g[i] - s1 - xh[i] * s2) + rv[i];
                dx[c] = round_out == OUT_TF32 ? round_tf32(o) : o;
                if (dX16 != nullptr) dX16[(long long)row * lddx16 + c] = __float2bfloat16_rn(o);   // shadow for the bf16 GEMMs
            }
        }
    }
    if (dgamma == nullptr) return;   // dgrad-only pass (path-length VJP)
#pragma unroll
    for (int i = 0; i < NPL; ++i) red[warp][i * 32 + lane] = dg[i];
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += LN_WARPS * 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) s += red[w][c];
        atomicAdd(dgamma + c, s);
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NPL; ++i) red[warp][i * 32 + lane] = db[i];
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += LN_WARPS * 32) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) s += red[w][c];
        atomicAdd(dbeta + c, s);
    }
}

}  // namespace

int launch_layernorm_fwd(const float* X, int ldx, const float* gamma, const float* beta, float* Y, int ldy,
                         float* mean, float* rstd, int M, int D, int round_out, cudaStream_t stream) {
    SCAT_REQUIRE(D >= 1 && D <= 1024, kErrUnsupported, "layernorm: D=%d not in [1,1024]", D);
    const int grid = ceil_div(M, LN_WARPS);
    if (D <= 256) SCAT_CHECK_CUDA(launch_k(layernorm_fwd_kernel<8>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, X, ldx, gamma, beta, Y, ldy, mean, rstd, M, D, round_out));
    else if (D <= 512) SCAT_CHECK_CUDA(launch_k(layernorm_fwd_kernel<16>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, X, ldx, gamma, beta, Y, ldy, mean, rstd, M, D, round_out));
    else SCAT_CHECK_CUDA(launch_k(layernorm_fwd_kernel<32>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, X, ldx, gamma, beta, Y, ldy, mean, rstd, M, D, round_out));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_layernorm_bwd(const float* dY, int lddy, const float* X, int ldx, const float* gamma, const float* mean,
                         const float* rstd, const float* resid, int ldr, float* dX, int lddx, float* dgamma,
                         float* dbeta, int M, int D, int round_out, cudaStream_t stream, int act_rows, void* dX16,
                         int lddx16) {
    SCAT_REQUIRE(D >= 1 && D <= 1024, kErrUnsupported, "layernorm bwd: D=%d not in [1,1024]", D);
    SCAT_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), kErrBadArg, "layernorm bwd: dgamma/dbeta must both be set or null");
    const int grid = min(ceil_div(M, 2 * LN_WARPS), 148);     // >= 2 rows per warp: halves the atomic tail
#define SCAT_LN_BWD(NPL) SCAT_CHECK_CUDA(launch_k(layernorm_bwd_kernel<NPL>, dim3(grid), dim3(LN_WARPS * 32), 0, stream,  \
        dY, lddy, X, ldx, gamma, mean, rstd, resid, ldr, dX, lddx, dgamma, dbeta, M, D, round_out, act_rows,  \
        reinterpret_cast<__nv_bfloat16*>(dX16), lddx16))
    if (D <= 256) SCAT_LN_BWD(8);
    else if (D <= 512) SCAT_LN_BWD(16);
    else if (D <= 800) SCAT_LN_BWD(25);    // d = 784: 25 elements per lane instead of 32 (255 registers + spills -> see -res-usage)
    else SCAT_LN_BWD(32);
#undef SCAT_LN_BWD
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace scat
