// Train-step tail: weak-perspective projection, 3D/2D losses, path-length statistic, and the closed-form
// gradient of the scalar loss w.r.t. pred_params (train.py:112-120,165-203).
//
//   cam = pred[:, :3] = (s,tx,ty);  J = pred[:, 3:66] as [21,3]
//   j2d = (J_xy + t_xy) * s * 112 + 112                                          :112-120,170-171
//   l_3d = mean((J - gt3d)^2) over B*63;  l_2d = mean(|j2d - gt2d|) over B*42    :188-192
//   len_b = sqrt(mean_t sum_yx pl[b,t]^2);  l_pl = mean_b (len_b - 0.01*mean(len))^2   (no gradient) :178-183
//   loss = w3d*l_3d + w2d*l_2d + 10*l_pl                                         :200-203
// Deterministic: fixed-order block reductions, no atomics.
#include "kernels.h"

namespace scat {
namespace {

constexpr int LOSS_THREADS = 256;

__global__ void __launch_bounds__(LOSS_THREADS)
rowsumsq_kernel(const float* __restrict__ X, int row_elems, float* __restrict__ out) {
    pdl_sync();
    __shared__ float part[LOSS_THREADS / 32];
    const float* x = X + (long long)blockIdx.x * row_elems;
    float s = 0.f;
    for (int i = threadIdx.x; i < row_elems; i += LOSS_THREADS) s = fmaf(x[i], x[i], s);
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < LOSS_THREADS / 32; ++i) t += part[i];
        out[blockIdx.x] = t;
    }
}

__device__ float block_sum(float v, float* scratch) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = 0.f;
    for (int i = 0; i < LOSS_THREADS / 32; ++i) t += scratch[i];
    return t;
}

__global__ void __launch_bounds__(LOSS_THREADS)
proj_loss_kernel(const float* __restrict__ pred, const float* __restrict__ labels, int ld_labels,
                 const float* __restrict__ pl_sumsq, int n_tokens, float w3d, float w2d, float grad_scale,
                 float* __restrict__ losses, float* __restrict__ g_pred, int B) {
    pdl_sync();
    __shared__ float scratch[LOSS_THREADS / 32];
    const int tid = threadIdx.x;
    const float inv3 = 1.0f / (63.0f * (float)B), inv2 = 1.0f / (42.0f * (float)B);
    // train.py:188-199: 105-wide rows are [63 3D | 42 2D]; wider rows (FreiHAND / HO-3D) are [61 pose | 63 3D | 42 2D]
    const int off3 = ld_labels == 105 ? 0 : 61;
    float a3 = 0.f, a2 = 0.f;
    // one thread per (sample, joint): 3D term for 3 coords, 2D term for 2 coords
    for (int e = tid; e < B * 21; e += LOSS_THREADS) {
        const int b = e / 21, t = e % 21;
        const float* p = pred + (long long)b * 66;
        const float* lb = labels + (long long)b * ld_labels + off3;
        const float s = p[0], tx = p[1], ty = p[2];
        const float jx = p[3 + 3 * t], jy = p[4 + 3 * t], jz = p[5 + 3 * t];
        const float dx = jx - lb[3 * t], dy = jy - lb[3 * t + 1], dz = jz - lb[3 * t + 2];
        a3 += dx * dx + dy * dy + dz * dz;
        const float ux = (jx + tx) * s * 112.0f + 112.0f - lb[63 + 2 * t];
        const float uy = (jy + ty) * s * 112.0f + 112.0f - lb[64 + 2 * t];
        a2 += fabsf(ux) + fabsf(uy);
        if (g_pred != nullptr) {
            const float sx = (ux > 0.f) - (ux < 0.f), sy = (uy > 0.f) - (uy < 0.f);
            const float c2 = w2d * 112.0f * inv2 * grad_scale;
            float* g = g_pred + (long long)b * 66;
            g[3 + 3 * t] = (w3d * 2.0f * inv3 * dx) * grad_scale + c2 * s * sx;
            g[4 + 3 * t] = (w3d * 2.0f * inv3 * dy) * grad_scale + c2 * s * sy;
            g[5 + 3 * t] = (w3d * 2.0f * inv3 * dz) * grad_scale;
        }
    }
    const float l3 = block_sum(a3, scratch) * inv3;
    const float l2 = block_sum(a2, scratch) * inv2;
    // camera gradients: one thread per sample, fixed order over joints
    if (g_pred != nullptr) {
        for (int b = tid; b < B; b += LOSS_THREADS) {
            const float* p = pred + (long long)b * 66;
            const float* lb = labels + (long long)b * ld_labels + off3;
            const float s = p[0], tx = p[1], ty = p[2];
            float gs = 0.f, gtx = 0.f, gty = 0.f;
            for (int t = 0; t < 21; ++t) {
                const float jx = p[3 + 3 * t], jy = p[4 + 3 * t];
                const float ux = (jx + tx) * s * 112.0f + 112.0f - lb[63 + 2 * t];
                const float uy = (jy + ty) * s * 112.0f + 112.0f - lb[64 + 2 * t];
                const float sx = (ux > 0.f) - (ux < 0.f), sy = (uy > 0.f) - (uy < 0.f);
                gs += (jx + tx) * sx + (jy + ty) * sy;
                gtx += sx;
                gty += sy;
            }
            const float c2 = w2d * 112.0f * inv2 * grad_scale;
            float* g = g_pred + (long long)b * 66;
            g[0] = c2 * gs;
            g[1] = c2 * s * gtx;
            g[2] = c2 * s * gty;
        }
    }
    float lpl = 0.f;
    if (pl_sumsq != nullptr) {
        float a = 0.f;
        for (int b = tid; b < B; b += LOSS_THREADS) a += sqrtf(pl_sumsq[b] / (float)n_tokens);
        const float pl_mean = 0.01f * (block_sum(a, scratch) / (float)B);
        float q = 0.f;
        for (int b = tid; b < B; b += LOSS_THREADS) {
            const float d = sqrtf(pl_sumsq[b] / (float)n_tokens) - pl_mean;
            q += d * d;
        }
        lpl = block_sum(q, scratch) / (float)B;
    }
    if (tid == 0) {
        losses[0] = w3d * l3 + w2d * l2 + (pl_sumsq != nullptr ? 10.0f * lpl : 0.f);
        losses[1] = l3;
        losses[2] = l2;
        losses[3] = lpl;
    }
}

__global__ void __launch_bounds__(LOSS_THREADS)
pl_loss_add_kernel(const float* __restrict__ pl_sumsq, int n_tokens, float* __restrict__ losses, int B) {
    pdl_sync();
    __shared__ float scratch[LOSS_THREADS / 32];
    const int tid = threadIdx.x;
    float a = 0.f;
    for (int b = tid; b < B; b += LOSS_THREADS) a += sqrtf(pl_sumsq[b] / (float)n_tokens);
    const float pl_mean = 0.01f * (block_sum(a, scratch) / (float)B);
    float q = 0.f;
    for (int b = tid; b < B; b += LOSS_THREADS) {
        const float d = sqrtf(pl_sumsq[b] / (float)n_tokens) - pl_mean;
        q += d * d;
    }
    const float lpl = block_sum(q, scratch) / (float)B;
    if (tid == 0) {
        losses[3] = lpl;
        losses[0] += 10.0f * lpl;
    }
}

}  // namespace

int launch_pl_loss_add(const float* pl_term, int pl_row_elems, int n_tokens, float* losses, float* pl_scratch, int B,
                       cudaStream_t stream) {
    SCAT_REQUIRE(pl_term && losses && pl_scratch && B > 0, kErrBadArg, "pl_loss_add: bad args");
    SCAT_CHECK_CUDA(launch_k(rowsumsq_kernel, dim3(B), dim3(LOSS_THREADS), 0, stream, pl_term, pl_row_elems, pl_scratch));
    SCAT_CHECK_LAUNCH();
    SCAT_CHECK_CUDA(launch_k(pl_loss_add_kernel, dim3(1), dim3(LOSS_THREADS), 0, stream, (const float*)pl_scratch, n_tokens, losses, B));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_proj_loss(const float* pred, const float* labels, int ld_labels, const float* pl_term, int pl_row_elems,
                     int n_tokens, float w3d, float w2d, float grad_scale, float* losses, float* g_pred,
                     float* pl_scratch, int B, cudaStream_t stream) {
    SCAT_REQUIRE(pred && labels && losses && B > 0, kErrBadArg, "proj_loss: bad args");
    // the reference slices by width (train.py:188-199) and its L1 loss needs exactly 42 2D columns behind the 3D block
    SCAT_REQUIRE(ld_labels == 105 || ld_labels == 166, kErrBadArg,
                 "proj_loss: label rows are 105 wide (63 3D + 42 2D) or 166 wide (61 pose + 63 3D + 42 2D), got %d", ld_labels);
    if (pl_term != nullptr) {
        SCAT_REQUIRE(pl_scratch != nullptr, kErrBadArg, "proj_loss: pl_scratch[B] required with pl_term");
        SCAT_CHECK_CUDA(launch_k(rowsumsq_kernel, dim3(B), dim3(LOSS_THREADS), 0, stream, pl_term, pl_row_elems, pl_scratch));
        SCAT_CHECK_LAUNCH();
    }
    SCAT_CHECK_CUDA(launch_k(proj_loss_kernel, dim3(1), dim3(LOSS_THREADS), 0, stream, pred, labels, ld_labels, pl_term ? pl_scratch : nullptr, n_tokens,
                                                     w3d, w2d, grad_scale, losses, g_pred, B));
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace scat
