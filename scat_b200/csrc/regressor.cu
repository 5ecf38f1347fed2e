// Autoregressive offset regressor of the head (hand_net.py:379-393) as one fused kernel, plus backward.
//
//   pred = mean_params (broadcast);  pred[3:] += feat_out                        :379-383
//   repeat `iteration` times:  pred += [main_feat | pred] Wr^T + br              :385-387
//   joints = pred[3:66] as [21,3];  joints -= joints[1]                          :389-393
//
// The product with main_feat is iteration-invariant, so it is computed once per sample
// (h = Wr[:, :F] main_feat + br, warp-level dot products) and the recurrence only applies the
// P x P block Wr[:, F:] held in shared memory.  Always fp32: the output error of the whole head is
// dominated by these K=1024 dot products.
// Also used with root_relative=0, P=61, F=1024 for the H3DWEncoder regressor (hand_net.py:53-57).
#include "kernels.h"

namespace scat {
namespace {

constexpr int REG_THREADS = 256;
constexpr int MAXP = 96;

__global__ void __launch_bounds__(REG_THREADS)
regressor_fwd_kernel(const float* __restrict__ main_feat, const float* __restrict__ feat_out,
                     const float* __restrict__ mean_params, const float* __restrict__ Wr, const float* __restrict__ br,
                     float* __restrict__ pred, float* __restrict__ states, int F, int P, int iteration,
                     int root_relative) {
    extern __shared__ float sm[];
    float* mf = sm;                    // [F]
    float* Wp = mf + F;                // [P][P+1]
    float* h = Wp + P * (P + 1);       // [P]
    float* p = h + P;                  // [P]
    float* pn = p + P;                 // [P]
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ldw = F + P;
    for (int k = tid; k < F; k += REG_THREADS) mf[k] = main_feat[(long long)b * F + k];
    for (int i = tid; i < P * P; i += REG_THREADS) Wp[(i / P) * (P + 1) + (i % P)] = Wr[(long long)(i / P) * ldw + F + (i % P)];
    if (tid < P) {
        float v = mean_params[tid];
        if (feat_out != nullptr && tid >= 3) v += feat_out[(long long)b * (P - 3) + tid - 3];
        p[tid] = v;
    }
    __syncthreads();
    for (int j = warp; j < P; j += REG_THREADS / 32) {
        const float* w = Wr + (long long)j * ldw;
        float s = 0.f;
        for (int k = lane; k < F; k += 32) s = fmaf(w[k], mf[k], s);
        s = warp_sum(s);
        if (lane == 0) h[j] = s + br[j];
    }
    __syncthreads();
    for (int it = 0; it < iteration; ++it) {
        if (tid < P) {
            if (states != nullptr) states[((long long)b * iteration + it) * P + tid] = p[tid];
            float s = h[tid];
            for (int k = 0; k < P; ++k) s = fmaf(Wp[tid * (P + 1) + k], p[k], s);
            pn[tid] = p[tid] + s;
        }
        __syncthreads();
        if (tid < P) p[tid] = pn[tid];
        __syncthreads();
    }
    if (tid < P) {
        float v = p[tid];
        if (root_relative && tid >= 3) v -= p[3 + 3 + (tid - 3) % 3];   // joint 1 is the root: exactly 0 afterwards
        pred[(long long)b * P + tid] = v;
    }
}

__global__ void __launch_bounds__(REG_THREADS)
regressor_bwd_kernel(const float* __restrict__ g_pred, const float* __restrict__ Wr, float* __restrict__ d_feat_out,
                     float* __restrict__ d_main_feat, float* __restrict__ gsum_out, float* __restrict__ gsteps, int F,
                     int P, int iteration, int root_relative) {
    extern __shared__ float sm[];
    float* Wp = sm;                    // [P][P+1]
    float* g = Wp + P * (P + 1);       // [P]
    float* gn = g + P;                 // [P]
    float* gsum = gn + P;              // [P]
    float* rootg = gsum + P;           // [3]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int ldw = F + P;
    for (int i = tid; i < P * P; i += REG_THREADS) Wp[(i / P) * (P + 1) + (i % P)] = Wr[(long long)(i / P) * ldw + F + (i % P)];
    if (tid < P) {
        g[tid] = g_pred[(long long)b * P + tid];
        gsum[tid] = 0.f;
    }
    __syncthreads();
    if (root_relative) {
        if (tid < 3) {
            float s = 0.f;
            for (int t = 0; t < (P - 3) / 3; ++t) s += g[3 + 3 * t + tid];
            rootg[tid] = s;
        }
        __syncthreads();
        if (tid < 3) g[3 + 3 + tid] -= rootg[tid];
        __syncthreads();
    }
    for (int it = iteration - 1; it >= 0; --it) {
        if (tid < P) {
            gsteps[((long long)b * iteration + it) * P + tid] = g[tid];
            gsum[tid] += g[tid];
            float s = g[tid];
            for (int j = 0; j < P; ++j) s = fmaf(Wp[j * (P + 1) + tid], g[j], s);   // (I + Wp^T) g
            gn[tid] = s;
        }
        __syncthreads();
        if (tid < P) g[tid] = gn[tid];
        __syncthreads();
    }
    if (tid < P) {
        gsum_out[(long long)b * P + tid] = gsum[tid];
        if (d_feat_out != nullptr && tid >= 3) d_feat_out[(long long)b * (P - 3) + tid - 3] = g[tid];
    }
    if (d_main_feat != nullptr) {
        for (int k = tid; k < F; k += REG_THREADS) {
            float s = 0.f;
            for (int j = 0; j < P; ++j) s = fmaf(Wr[(long long)j * ldw + k], gsum[j], s);
            d_main_feat[(long long)b * F + k] = s;
        }
    }
}

}  // namespace

int launch_regressor_fwd(const float* main_feat, const float* feat_out, const float* mean_params, const float* Wr,
                         const float* br, float* pred, float* states, int B, int F, int P, int iteration,
                         int root_relative, cudaStream_t stream) {
    SCAT_REQUIRE(P >= 4 && P <= MAXP && F >= 1 && F <= 4096, kErrUnsupported, "regressor: P=%d F=%d out of range", P, F);
    SCAT_REQUIRE(!root_relative || (P - 3) % 3 == 0, kErrBadArg, "regressor: root_relative needs P=3+3k");
    const size_t smem = sizeof(float) * ((size_t)F + (size_t)P * (P + 1) + 3 * (size_t)P);
    if (smem > 48 * 1024)
        SCAT_CHECK_CUDA(cudaFuncSetAttribute(regressor_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    regressor_fwd_kernel<<<B, REG_THREADS, smem, stream>>>(main_feat, feat_out, mean_params, Wr, br, pred, states, F, P,
                                                          iteration, root_relative);
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_regressor_bwd(const float* g_pred, const float* Wr, float* d_feat_out, float* d_main_feat, float* gsum,
                         float* gsteps, int B, int F, int P, int iteration, int root_relative, cudaStream_t stream) {
    SCAT_REQUIRE(P >= 4 && P <= MAXP, kErrUnsupported, "regressor bwd: P=%d out of range", P);
    SCAT_REQUIRE(gsum && gsteps, kErrBadArg, "regressor bwd: gsum/gsteps scratch required");
    const size_t smem = sizeof(float) * ((size_t)P * (P + 1) + 3 * (size_t)P + 4);
    regressor_bwd_kernel<<<B, REG_THREADS, smem, stream>>>(g_pred, Wr, d_feat_out, d_main_feat, gsum, gsteps, F, P,
                                                          iteration, root_relative);
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace scat
