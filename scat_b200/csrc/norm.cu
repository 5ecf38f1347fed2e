// LayerNorm forward / backward over token rows (vision_transformer.py:23,26: nn.LayerNorm(dim), eps=1e-5,
// affine).  One warp per row, the row lives in registers; two-pass mean/variance like ATen's CPU kernel.
#include "kernels.h"

namespace scat {
namespace {

constexpr float kEps = 1e-5f;
constexpr int LN_WARPS = 8;

template <int NPL, bool RESID>  // elements per lane (row length <= 32*NPL); RESID: Y = LN(X) + resid (coarse variant)
__global__ void __launch_bounds__(LN_WARPS * 32, 2)
layernorm_fwd_kernel(const float* __restrict__ X, int ldx, const float* __restrict__ gamma,
                     const float* __restrict__ beta, float* __restrict__ Y, int ldy, float* __restrict__ mean_out,
                     float* __restrict__ rstd_out, int M, int D, int round_out, const float* __restrict__ resid, int ldr) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * LN_WARPS + (threadIdx.x >> 5);
    if (row >= M) return;
    const float* x = X + (long long)row * ldx;
    const float* rr = RESID ? resid + (long long)row * ldr : nullptr;
    // every global read of the row (x, gamma, beta, residual) is issued before the first store: Y may alias any of them as
    // far as the compiler knows (store_out also writes through a bf16 pointer), and a load placed after a store waits one L2
    // round trip per element -- ncu had this kernel at 17 us for 6 MB with 25 long-scoreboard stalls per issue
    float v[NPL], gm[NPL], bt[NPL], rv[RESID ? NPL : 1];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
        const int c = i * 32 + lane;
        v[i] = c < D ? __ldg(x + c) : 0.f;
        gm[i] = c < D ? __ldg(gamma + c) : 0.f;
        bt[i] = c < D ? __ldg(beta + c) : 0.f;
        if (RESID) rv[i] = c < D ? __ldg(rr + c) : 0.f;             // x = pren(x1) + x (vision_transformer_attn.py:108)
    }
#pragma unroll
    for (int i = 0; i < NPL; ++i) s += v[i];
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
            float o = (v[i] - mean) * rstd * gm[i] + bt[i];
            if (RESID) o += rv[i];
            store_out(Y, (long long)row * ldy + c, o, round_out);
        }
    }
    if (lane == 0) {
        mean_out[row] = mean;
        rstd_out[row] = rstd;
    }
}

// PARAMS: also accumulate dgamma / dbeta (the stand-alone operator).  The head runs the data gradient alone on its
// critical path (PARAMS = false: no per-lane accumulators, half the registers, two CTAs per SM) and the parameter
// gradients as column sums on the side stream (layernorm_param_grad_kernel).
template <int NPL, bool PARAMS>
__global__ void __launch_bounds__(LN_WARPS * 32, PARAMS ? 1 : 2)
layernorm_bwd_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ X, int ldx,
                     const float* __restrict__ gamma, const float* __restrict__ mean, const float* __restrict__ rstd,
                     const float* __restrict__ resid, int ldr, float* __restrict__ dX, int lddx,
                     float* __restrict__ dgamma, float* __restrict__ dbeta, int M, int D, int round_out, int act_rows,
                     __nv_bfloat16* __restrict__ dX16, int lddx16) {
    pdl_sync();
    __shared__ float red[PARAMS ? LN_WARPS : 1][PARAMS ? 32 * NPL : 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float dg[PARAMS ? NPL : 1], db[PARAMS ? NPL : 1];
#pragma unroll
    for (int i = 0; i < (PARAMS ? NPL : 1); ++i) {
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
            if (PARAMS && real) {
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
    if constexpr (PARAMS) {
        if (dgamma == nullptr) return;
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
}

// dgamma[c] += sum_m dY[m,c] * (X[m,c] - mean[m]) * rstd[m],  dbeta[c] += sum_m dY[m,c]   (both zero on entry).
// lane = column, the 8 warps of a block stride the rows of the block's row slice; slices combine with atomics.
constexpr int PG_SLICES = 12;
__global__ void __launch_bounds__(LN_WARPS * 32)
layernorm_param_grad_kernel(const float* __restrict__ dY, int lddy, const float* __restrict__ X, int ldx,
                            const float* __restrict__ mean, const float* __restrict__ rstd, float* __restrict__ dgamma,
                            float* __restrict__ dbeta, int M, int D, const float* __restrict__ E, int lde,
                            float* __restrict__ esum) {
    pdl_sync();
    __shared__ float red[3][LN_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const int rows_per = (M + gridDim.y - 1) / gridDim.y;
    const int r0 = blockIdx.y * rows_per, r1 = min(M, r0 + rows_per);
    float dg = 0.f, db = 0.f, de = 0.f;
    if (c < D) {
        if (E != nullptr) {        // a bias gradient riding along: esum[c] += sum_m E[m,c]
#pragma unroll 4
            for (int r = r0 + warp; r < r1; r += LN_WARPS) {
                const float dv = __ldg(dY + (long long)r * lddy + c);
                const float xh = (__ldg(X + (long long)r * ldx + c) - __ldg(mean + r)) * __ldg(rstd + r);
                dg = fmaf(dv, xh, dg);
                db += dv;
                de += __ldg(E + (long long)r * lde + c);
            }
        } else {
#pragma unroll 4
            for (int r = r0 + warp; r < r1; r += LN_WARPS) {
                const float dv = __ldg(dY + (long long)r * lddy + c);
                const float xh = (__ldg(X + (long long)r * ldx + c) - __ldg(mean + r)) * __ldg(rstd + r);
                dg = fmaf(dv, xh, dg);
                db += dv;
            }
        }
    }
    red[0][warp][lane] = dg;
    red[1][warp][lane] = db;
    red[2][warp][lane] = de;
    __syncthreads();
    if (warp < (E != nullptr ? 3 : 2) && c < D) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < LN_WARPS; ++w) s += red[warp][w][lane];
        atomicAdd((warp == 0 ? dgamma : warp == 1 ? dbeta : esum) + c, s);
    }
}

}  // namespace

int launch_layernorm_param_grads(const float* dY, int lddy, const float* X, int ldx, const float* mean, const float* rstd,
                                 float* dgamma, float* dbeta, int M, int D, cudaStream_t stream, const float* E, int lde,
                                 float* esum) {
    SCAT_REQUIRE(dY && X && mean && rstd && dgamma && dbeta && M > 0 && D > 0, kErrBadArg, "layernorm param grads: bad args");
    SCAT_REQUIRE((E == nullptr) == (esum == nullptr), kErrBadArg, "layernorm param grads: E and esum go together");
    SCAT_CHECK_CUDA(launch_k(layernorm_param_grad_kernel, dim3(ceil_div(D, 32), min(PG_SLICES, ceil_div(M, 64))), dim3(LN_WARPS * 32), 0,
                             stream, dY, lddy, X, ldx, mean, rstd, dgamma, dbeta, M, D, E, lde, esum));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_layernorm_fwd(const float* X, int ldx, const float* gamma, const float* beta, float* Y, int ldy,
                         float* mean, float* rstd, int M, int D, int round_out, cudaStream_t stream, const float* resid,
                         int ldr) {
    SCAT_REQUIRE(D >= 1 && D <= 1024, kErrUnsupported, "layernorm: D=%d not in [1,1024]", D);
    const int grid = ceil_div(M, LN_WARPS);
#define SCAT_LN_FWD(NPL)                                                                                                      \
    do {                                                                                                                      \
        if (resid != nullptr)                                                                                                 \
            SCAT_CHECK_CUDA(launch_k(layernorm_fwd_kernel<NPL, true>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, X, ldx, gamma, \
                                     beta, Y, ldy, mean, rstd, M, D, round_out, resid, ldr));                                   \
        else                                                                                                                  \
            SCAT_CHECK_CUDA(launch_k(layernorm_fwd_kernel<NPL, false>, dim3(grid), dim3(LN_WARPS * 32), 0, stream, X, ldx, gamma, \
                                     beta, Y, ldy, mean, rstd, M, D, round_out, resid, ldr));                                   \
    } while (0)
    if (D <= 256) SCAT_LN_FWD(8);
    else if (D <= 512) SCAT_LN_FWD(16);
    else if (D <= 800) SCAT_LN_FWD(25);
    else SCAT_LN_FWD(32);
#undef SCAT_LN_FWD
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_layernorm_bwd(const float* dY, int lddy, const float* X, int ldx, const float* gamma, const float* mean,
                         const float* rstd, const float* resid, int ldr, float* dX, int lddx, float* dgamma,
                         float* dbeta, int M, int D, int round_out, cudaStream_t stream, int act_rows, void* dX16,
                         int lddx16) {
    SCAT_REQUIRE(D >= 1 && D <= 1024, kErrUnsupported, "layernorm bwd: D=%d not in [1,1024]", D);
    SCAT_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), kErrBadArg, "layernorm bwd: dgamma/dbeta must both be set or null");
    const bool params = dgamma != nullptr;
    // with parameter gradients: >= 2 rows per warp halves the atomic tail; without: one row per warp, two CTAs per SM
    const int grid = params ? min(ceil_div(M, 2 * LN_WARPS), 148) : min(ceil_div(M, LN_WARPS), 2 * 148);
#define SCAT_LN_BWD(NPL, PAR) SCAT_CHECK_CUDA(launch_k(layernorm_bwd_kernel<NPL, PAR>, dim3(grid), dim3(LN_WARPS * 32), 0, stream,  \
        dY, lddy, X, ldx, gamma, mean, rstd, resid, ldr, dX, lddx, dgamma, dbeta, M, D, round_out, act_rows,  \
        reinterpret_cast<__nv_bfloat16*>(dX16), lddx16))
#define SCAT_LN_BWD2(NPL) do { if (params) SCAT_LN_BWD(NPL, true); else SCAT_LN_BWD(NPL, false); } while (0)
    if (D <= 256) SCAT_LN_BWD2(8);
    else if (D <= 512) SCAT_LN_BWD2(16);
    else if (D <= 800) SCAT_LN_BWD2(25);   // d = 784: 25 elements per lane instead of 32
    else SCAT_LN_BWD2(32);
#undef SCAT_LN_BWD2
#undef SCAT_LN_BWD
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace scat
