// Autoregressive offset regressor of the head (hand_net.py:379-393), fused, plus its backward.
//
//   pred = mean_params (broadcast);  pred[3:] += feat_out                        :379-383
//   repeat `iteration` times:  pred += [main_feat | pred] Wr^T + br              :385-387
//   joints = pred[3:66] as [21,3];  joints -= joints[1]                          :389-393
//
// The product with main_feat is iteration-invariant, so it is hoisted: h = Wr[:, :F] main_feat + br is
// computed once (kernel 1: warp-level dot products, 8 samples share every weight row from registers), and the
// recurrence (kernel 2: one CTA per sample) only applies the P x P block Wr[:, F:] held in shared memory.
// Always fp32: the output error of the whole head is dominated by these K=1024 dot products.
// Also used with root_relative=0, P=61, F=1024 for the H3DWEncoder regressor (hand_net.py:53-57).
#include "kernels.h"

namespace scat {
namespace {

constexpr int MAXP = 96;
constexpr int HS = 8;              // samples per CTA in the hoist kernel
constexpr int HJ = 8;              // weight rows per CTA (one per warp)

// h[b,j] = br[j] + sum_k Wr[j,k] mf[b,k],  k < F.   grid (ceil(B/HS), ceil(P/HJ)), 256 threads
__global__ void __launch_bounds__(256)
regressor_hoist_kernel(const float* __restrict__ main_feat, const float* __restrict__ Wr, const float* __restrict__ br,
                       float* __restrict__ h, int B, int F, int P) {
    pdl_sync();
    extern __shared__ __align__(16) float mfs[];   // [HS][F]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b0 = blockIdx.x * HS, j = blockIdx.y * HJ + warp;
    const int ns = min(HS, B - b0);
    for (int i = tid; i < HS * (F >> 1); i += 256) {
        const int s = i / (F >> 1), k2 = i % (F >> 1);
        reinterpret_cast<float2*>(mfs)[i] = s < ns ? __ldg(reinterpret_cast<const float2*>(main_feat + (long long)(b0 + s) * F) + k2)
                                                   : make_float2(0.f, 0.f);
    }
    __syncthreads();
    if (j >= P) return;
    const int ldw = F + P;
    float acc[HS];
#pragma unroll
    for (int s = 0; s < HS; ++s) acc[s] = 0.f;
    if ((ldw & 1) == 0) {                   // rows 8-byte aligned (EncoderTransformer: 1090): float2 loads
        const float2* w = reinterpret_cast<const float2*>(Wr + (long long)j * ldw);
        for (int k2 = lane; k2 < (F >> 1); k2 += 32) {
            const float2 wv = __ldg(w + k2);
#pragma unroll
            for (int s = 0; s < HS; ++s) {
                const float2 m = reinterpret_cast<const float2*>(mfs + s * F)[k2];
                acc[s] = fmaf(wv.x, m.x, acc[s]);
                acc[s] = fmaf(wv.y, m.y, acc[s]);
            }
        }
    } else {                                // odd leading dimension (H3DWEncoder: 1085): scalar loads
        const float* w = Wr + (long long)j * ldw;
        for (int k = lane; k < F; k += 32) {
            const float wv = __ldg(w + k);
#pragma unroll
            for (int s = 0; s < HS; ++s) acc[s] = fmaf(wv, mfs[s * F + k], acc[s]);
        }
    }
#pragma unroll
    for (int s = 0; s < HS; ++s) acc[s] = warp_sum(acc[s]);
    if (lane == 0) {
        const float bj = br[j];
        for (int s = 0; s < ns; ++s) h[(long long)(b0 + s) * P + j] = acc[s] + bj;
    }
}

// recurrence + root-relative step; one CTA per sample, 128 threads
__global__ void __launch_bounds__(128)
regressor_iter_kernel(const float* __restrict__ h_all, const float* __restrict__ feat_out,
                      const float* __restrict__ mean_params, const float* __restrict__ Wr, float* __restrict__ pred,
                      float* __restrict__ states, int F, int P, int iteration, int root_relative) {
    pdl_sync();
    extern __shared__ float sm[];
    float* Wp = sm;                    // [P][P+1]
    float* h = Wp + P * (P + 1);       // [P]
    float* p = h + P;                  // [P]
    float* pn = p + P;                 // [P]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int ldw = F + P;
    for (int i = tid; i < P * P; i += 128) Wp[(i / P) * (P + 1) + (i % P)] = __ldg(Wr + (long long)(i / P) * ldw + F + (i % P));
    if (tid < P) {
        float v = mean_params[tid];
        if (feat_out != nullptr && tid >= 3) v += feat_out[(long long)b * (P - 3) + tid - 3];
        p[tid] = v;
        h[tid] = h_all[(long long)b * P + tid];
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

// reverse recurrence: g_pred -> gsum (sum over steps of the per-step output gradient), gsteps, d feat_out
__global__ void __launch_bounds__(128)
regressor_iter_bwd_kernel(const float* __restrict__ g_pred, const float* __restrict__ Wr, float* __restrict__ d_feat_out,
                          float* __restrict__ gsum_out, float* __restrict__ gsteps, int F, int P, int iteration,
                          int root_relative) {
    pdl_sync();
    extern __shared__ float sm[];
    float* Wp = sm;                    // [P][P+1]
    float* g = Wp + P * (P + 1);       // [P]
    float* gn = g + P;                 // [P]
    float* rootg = gn + P;             // [3]
    const int b = blockIdx.x, tid = threadIdx.x;
    const int ldw = F + P;
    for (int i = tid; i < P * P; i += 128) Wp[(i / P) * (P + 1) + (i % P)] = __ldg(Wr + (long long)(i / P) * ldw + F + (i % P));
    float gsum = 0.f;
    if (tid < P) g[tid] = g_pred[(long long)b * P + tid];
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
            gsum += g[tid];
            float s = g[tid];
            for (int j = 0; j < P; ++j) s = fmaf(Wp[j * (P + 1) + tid], g[j], s);   // (I + Wp^T) g
            gn[tid] = s;
        }
        __syncthreads();
        if (tid < P) g[tid] = gn[tid];
        __syncthreads();
    }
    if (tid < P) {
        gsum_out[(long long)b * P + tid] = gsum;
        if (d_feat_out != nullptr && tid >= 3) d_feat_out[(long long)b * (P - 3) + tid - 3] = g[tid];
    }
}

// The fused train step's tail (csrc/head.cu): forward recurrence, root-relative step, the closed-form gradient of the
// scalar loss w.r.t. pred_params (train.py:112-120,188-203, same arithmetic as proj_loss_kernel in loss.cu) and the
// reverse recurrence, in ONE kernel per sample -- the loss VALUES need a reduction over the batch, the gradient does
// not (the batch only enters through the constants 1/(63B), 1/(42B)), so the values are computed off the critical path.
// Replaces regressor_iter_kernel -> proj_loss_kernel -> regressor_iter_bwd_kernel (+ the ones-cotangent fill) there.
__global__ void __launch_bounds__(128)
regressor_train_kernel(const float* __restrict__ h_all, const float* __restrict__ feat_out,
                       const float* __restrict__ mean_params, const float* __restrict__ Wr, const float* __restrict__ labels,
                       int ld_labels, float w3d, float w2d, float grad_scale, float* __restrict__ pred,
                       float* __restrict__ states, float* __restrict__ d_feat_out, float* __restrict__ ones_out,
                       float* __restrict__ gsum_out, float* __restrict__ gsteps, int B, int F, int iteration) {
    pdl_sync();
    constexpr int P = 66, NJT = 21;
    extern __shared__ float sm[];
    float* Wp = sm;                    // [P][P+1]
    float* h = Wp + P * (P + 1);       // [P]
    float* p = h + P;                  // [P]
    float* pn = p + P;                 // [P]
    float* cam = pn + P;               // [3][NJT] per-joint camera-gradient terms
    const int b = blockIdx.x, tid = threadIdx.x;
    const int ldw = F + P;
    for (int i = tid; i < P * P; i += 128) Wp[(i / P) * (P + 1) + (i % P)] = __ldg(Wr + (long long)(i / P) * ldw + F + (i % P));
    if (tid < P) {
        float v = mean_params[tid];
        if (tid >= 3) v += feat_out[(long long)b * (P - 3) + tid - 3];
        p[tid] = v;
        h[tid] = h_all[(long long)b * P + tid];
    }
    if (ones_out != nullptr && tid < P - 3) ones_out[(long long)b * (P - 3) + tid] = 1.0f;   // path-length cotangent (hand_net.py:396)
    __syncthreads();
    // ---- forward recurrence (hand_net.py:385-387) ----
    for (int it = 0; it < iteration; ++it) {
        if (tid < P) {
            states[((long long)b * iteration + it) * P + tid] = p[tid];
            float s = h[tid];
            for (int k = 0; k < P; ++k) s = fmaf(Wp[tid * (P + 1) + k], p[k], s);
            pn[tid] = p[tid] + s;
        }
        __syncthreads();
        if (tid < P) p[tid] = pn[tid];
        __syncthreads();
    }
    // ---- root-relative joints (hand_net.py:389-393) ----
    if (tid < P) {
        float v = p[tid];
        if (tid >= 3) v -= p[3 + 3 + (tid - 3) % 3];
        pred[(long long)b * P + tid] = v;
        pn[tid] = v;
    }
    __syncthreads();
    // ---- d loss / d pred_params, closed form (loss.cu: proj_loss_kernel) into p[] ----
    const float inv3 = 1.0f / (63.0f * (float)B), inv2 = 1.0f / (42.0f * (float)B);
    const float c2 = w2d * 112.0f * inv2 * grad_scale;
    const float* lb = labels + (long long)b * ld_labels + (ld_labels == 105 ? 0 : 61);   // train.py:188-199
    if (tid < NJT) {
        const int t = tid;
        const float s = pn[0], tx = pn[1], ty = pn[2];
        const float jx = pn[3 + 3 * t], jy = pn[4 + 3 * t], jz = pn[5 + 3 * t];
        const float dx = jx - lb[3 * t], dy = jy - lb[3 * t + 1], dz = jz - lb[3 * t + 2];
        const float ux = (jx + tx) * s * 112.0f + 112.0f - lb[63 + 2 * t];
        const float uy = (jy + ty) * s * 112.0f + 112.0f - lb[64 + 2 * t];
        const float sx = (ux > 0.f) - (ux < 0.f), sy = (uy > 0.f) - (uy < 0.f);
        p[3 + 3 * t] = (w3d * 2.0f * inv3 * dx) * grad_scale + c2 * s * sx;
        p[4 + 3 * t] = (w3d * 2.0f * inv3 * dy) * grad_scale + c2 * s * sy;
        p[5 + 3 * t] = (w3d * 2.0f * inv3 * dz) * grad_scale;
        cam[t] = (jx + tx) * sx + (jy + ty) * sy;
        cam[NJT + t] = sx;
        cam[2 * NJT + t] = sy;
    }
    __syncthreads();
    if (tid < 3) {                     // camera gradients: fixed order over the joints, like proj_loss_kernel
        float a = 0.f;
        for (int t = 0; t < NJT; ++t) a += cam[tid * NJT + t];
        p[tid] = tid == 0 ? c2 * a : c2 * pn[0] * a;
    }
    __syncthreads();
    // ---- backward through the root-relative step and the recurrence (regressor_iter_bwd_kernel) ----
    float* g = p;
    float* gn = pn;
    if (tid < 3) {
        float s = 0.f;
        for (int t = 0; t < NJT; ++t) s += g[3 + 3 * t + tid];
        cam[tid] = s;
    }
    __syncthreads();
    if (tid < 3) g[3 + 3 + tid] -= cam[tid];
    __syncthreads();
    float gsum = 0.f;
    for (int it = iteration - 1; it >= 0; --it) {
        if (tid < P) {
            gsteps[((long long)b * iteration + it) * P + tid] = g[tid];
            gsum += g[tid];
            float s = g[tid];
            for (int j = 0; j < P; ++j) s = fmaf(Wp[j * (P + 1) + tid], g[j], s);   // (I + Wp^T) g
            gn[tid] = s;
        }
        __syncthreads();
        if (tid < P) g[tid] = gn[tid];
        __syncthreads();
    }
    if (tid < P) {
        gsum_out[(long long)b * P + tid] = gsum;
        if (tid >= 3) d_feat_out[(long long)b * (P - 3) + tid - 3] = g[tid];
    }
}

// Everything the regressor contributes to the gradient bucket and to main_feat, in ONE launch (three FFMA GEMMs with a
// 66-wide dimension and a column sum before: ~110 us of side-stream time for 30 MFLOP).  128 threads per block:
//   blockIdx.y = 0   d main_feat[b, f] = sum_p gsum[b,p] Wr[p, f]                   thread = f, blockIdx.z = sample slice
//   blockIdx.y = 1   dWr[p, f]        = sum_b gsum[b,p] main_feat[b, f]   (f < F)   thread = f, blockIdx.z = slice of p
//   blockIdx.y = 2   dWr[p, F + q]    = sum_{b,it} gsteps[b,it,p] states[b,it,q];  dbr[p] = sum_b gsum[b,p]   (z = 0 only)
// Every output element is written exactly once (no atomics, no dependence on the bucket having been zeroed).
constexpr int PG_Z = 6, PG_SL = 16;      // slices; samples (y = 0) or rows p (y = 1) per slice, PG_Z * PG_SL >= 96
__global__ void __launch_bounds__(128)
regressor_param_grads_kernel(const float* __restrict__ gsum, const float* __restrict__ gsteps, const float* __restrict__ states,
                             const float* __restrict__ main_feat, const float* __restrict__ Wr, float* __restrict__ d_main_feat,
                             float* __restrict__ dWr, float* __restrict__ dbr, int B, int F, int P, int iteration) {
    pdl_sync();
    const int tid = threadIdx.x, ldw = F + P;
    const int f = blockIdx.x * 128 + tid;
    if (blockIdx.y == 0) {
        if (d_main_feat == nullptr) return;
        __shared__ float gs[PG_SL][MAXP];
        for (int b0 = blockIdx.z * PG_SL; b0 < B; b0 += PG_Z * PG_SL) {         // (B <= 96: one pass)
            const int nb = min(PG_SL, B - b0);
            __syncthreads();
            for (int i = tid; i < nb * P; i += 128) gs[i / P][i % P] = gsum[(long long)(b0 + i / P) * P + i % P];
            __syncthreads();
            if (f < F) {
                float acc[PG_SL];
#pragma unroll
                for (int b = 0; b < PG_SL; ++b) acc[b] = 0.f;
                for (int p = 0; p < P; ++p) {
                    const float w = __ldg(Wr + (long long)p * ldw + f);
#pragma unroll
                    for (int b = 0; b < PG_SL; ++b) acc[b] = fmaf(gs[b][p], w, acc[b]);     // rows >= nb hold stale finite data
                }
#pragma unroll
                for (int b = 0; b < PG_SL; ++b)
                    if (b < nb) d_main_feat[(long long)(b0 + b) * F + f] = acc[b];
            }
        }
    } else if (blockIdx.y == 1) {
        const int p0 = blockIdx.z * PG_SL;
        if (p0 >= P || f >= F) return;
        const int np = min(PG_SL, P - p0);
        float acc[PG_SL];
#pragma unroll
        for (int i = 0; i < PG_SL; ++i) acc[i] = 0.f;
        for (int b = 0; b < B; ++b) {
            const float m = __ldg(main_feat + (long long)b * F + f);
            const float* g = gsum + (long long)b * P + p0;
#pragma unroll
            for (int i = 0; i < PG_SL; ++i)
                if (i < np) acc[i] = fmaf(__ldg(g + i), m, acc[i]);
        }
#pragma unroll
        for (int i = 0; i < PG_SL; ++i)
            if (i < np) dWr[(long long)(p0 + i) * ldw + f] = acc[i];
    } else {
        if (blockIdx.z != 0) return;
        const int i = blockIdx.x * 128 + tid;
        if (i < P * P) {
            const int p = i / P, q = i - p * P;
            float a = 0.f;
            for (int r = 0; r < B * iteration; ++r) a = fmaf(__ldg(gsteps + (long long)r * P + p), __ldg(states + (long long)r * P + q), a);
            dWr[(long long)p * ldw + F + q] = a;
        } else if (i < P * P + P) {
            const int p = i - P * P;
            float a = 0.f;
            for (int b = 0; b < B; ++b) a += __ldg(gsum + (long long)b * P + p);
            dbr[p] = a;
        }
    }
}

// the last feed-forward of the head narrows to THREE outputs per token (vision_transformer.py:37-42): its second Linear
// and that Linear's data gradient are far too thin for a tiled GEMM (N = 3 / K = 3); always fp32.
//   Y[m, j] = b2[j] + sum_k H[m,k] W2[j,k]          one warp per row
__global__ void __launch_bounds__(256)
ff_out3_fwd_kernel(const float* __restrict__ H, int ldh, const float* __restrict__ W2, const float* __restrict__ b2,
                   float* __restrict__ Y, int M, int K) {
    pdl_sync();
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= M) return;
    const float* h = H + (long long)row * ldh;
    float a0 = 0.f, a1 = 0.f, a2 = 0.f;
    for (int k = lane; k < K; k += 32) {
        const float v = __ldg(h + k);
        a0 = fmaf(v, __ldg(W2 + k), a0);
        a1 = fmaf(v, __ldg(W2 + K + k), a1);
        a2 = fmaf(v, __ldg(W2 + 2 * K + k), a2);
    }
    a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2);
    if (lane < 3) Y[(long long)row * 3 + lane] = (lane == 0 ? a0 : lane == 1 ? a1 : a2) + __ldg(b2 + lane);
}
//   dZ[m, n] = (sum_{j<3} dY[m,j] W2[j,n]) * gelu'(Z[m % act_rows, n])      thread per element; the pad columns n >= N of a row
//   (lddz = N padded to 8 floats, head.cu: padp) are written as zeros so that the row can feed a tensor-core GEMM
__global__ void __launch_bounds__(256)
ff_out3_bwd_kernel(const float* __restrict__ dY, const float* __restrict__ W2, const float* __restrict__ Z, int ldz,
                   float* __restrict__ dZ, int lddz, int MR, int N, int act_rows, float* __restrict__ dZs, int z_is_grad) {
    pdl_sync();
    const int np = (N + 7) & ~7;
    const long long total = (long long)MR * np;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / np), n = (int)(i - (long long)m * np);
        float v = 0.f;
        if (n < N) {
            const float* dy = dY + (long long)m * 3;
            const float t = fmaf(__ldg(dy), __ldg(W2 + n), fmaf(__ldg(dy + 1), __ldg(W2 + N + n), __ldg(dy + 2) * __ldg(W2 + 2 * N + n)));
            const int zr = act_rows > 0 ? m % act_rows : m;
            const float z = __ldg(Z + (long long)zr * ldz + n);
            v = t * (z_is_grad ? z : gelu_erf_grad(z));       // z_is_grad: the forward saved gelu'(z) (GemmArgs::gelu_saves_grad)
        }
        if (n < lddz) dZ[(long long)m * lddz + n] = v;
        if (dZs != nullptr) {           // [hi | lo | hi], thirds of np columns: first operand of the 3xTF32 dgrad GEMM
            const float hi = round_tf32(v), lo = round_tf32(v - hi);
            float* d = dZs + (long long)m * 3 * np + n;
            d[0] = hi; d[np] = lo; d[2 * np] = hi;
        }
    }
}

}  // namespace

int launch_regressor_param_grads(const float* gsum, const float* gsteps, const float* states, const float* main_feat,
                                 const float* Wr, float* d_main_feat, float* dWr, float* dbr, int B, int F, int P, int iteration,
                                 cudaStream_t stream) {
    SCAT_REQUIRE(gsum && gsteps && states && main_feat && Wr && dWr && dbr, kErrBadArg, "regressor param grads: null argument");
    SCAT_REQUIRE(P >= 4 && P <= MAXP && P <= PG_Z * PG_SL && F >= 1, kErrUnsupported, "regressor param grads: P=%d F=%d", P, F);
    const int gx = max(ceil_div(F, 128), ceil_div(P * P + P, 128));
    SCAT_CHECK_CUDA(launch_k(regressor_param_grads_kernel, dim3(gx, 3, PG_Z), dim3(128), 0, stream, gsum, gsteps, states, main_feat, Wr,
                             d_main_feat, dWr, dbr, B, F, P, iteration));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_ff_out3_fwd(const float* H, int ldh, const float* W2, const float* b2, float* Y, int M, int K, cudaStream_t stream) {
    SCAT_CHECK_CUDA(launch_k(ff_out3_fwd_kernel, dim3(ceil_div(M, 8)), dim3(256), 0, stream, H, ldh, W2, b2, Y, M, K));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_ff_out3_bwd(const float* dY, const float* W2, const float* Z, int ldz, float* dZ, int lddz, int MR, int N, int act_rows,
                       float* dZs, cudaStream_t stream, int z_is_grad) {
    const long long total = (long long)MR * ((N + 7) & ~7);
    const int grid = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    SCAT_CHECK_CUDA(launch_k(ff_out3_bwd_kernel, dim3(grid), dim3(256), 0, stream, dY, W2, Z, ldz, dZ, lddz, MR, N, act_rows, dZs, z_is_grad));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_regressor_train(const float* feat_out, const float* mean_params, const float* Wr, const float* h_scratch,
                           const float* labels, int ld_labels, float w3d, float w2d, float grad_scale, float* pred,
                           float* states, float* d_feat_out, float* ones_out, float* gsum, float* gsteps, int B, int F, int P,
                           int iteration, cudaStream_t stream) {
    SCAT_REQUIRE(P == 66 && iteration >= 1, kErrUnsupported, "regressor train kernel: P=66, iteration>=1 (got %d, %d)", P, iteration);
    SCAT_REQUIRE(ld_labels == 105 || ld_labels == 166, kErrBadArg,
                 "label rows are 105 wide (63 3D + 42 2D) or 166 wide (61 pose + 63 3D + 42 2D), got %d", ld_labels);
    SCAT_REQUIRE(feat_out && mean_params && Wr && h_scratch && labels && pred && states && d_feat_out && gsum && gsteps, kErrBadArg,
                 "regressor train kernel: null argument");
    const size_t smem = sizeof(float) * ((size_t)P * (P + 1) + 3 * (size_t)P + 3 * 21);
    SCAT_CHECK_CUDA(launch_k(regressor_train_kernel, dim3(B), dim3(128), smem, stream, h_scratch, feat_out, mean_params, Wr, labels,
                             ld_labels, w3d, w2d, grad_scale, pred, states, d_feat_out, ones_out, gsum, gsteps, B, F, iteration));
    SCAT_CHECK_LAUNCH();
    return 0;
}

// the iteration-invariant part h = main_feat Wr[:, :F]^T + br: depends on the backbone feature only, so a caller may
// run it on another stream while the transformer computes feat_out
int launch_regressor_hoist(const float* main_feat, const float* Wr, const float* br, float* h_scratch, int B, int F, int P,
                           cudaStream_t stream) {
    SCAT_REQUIRE(P >= 4 && P <= MAXP && F >= 2 && F <= 4096 && (F & 1) == 0, kErrUnsupported,
                 "regressor: P=%d F=%d out of range (F even, P<=96, F<=4096)", P, F);
    SCAT_REQUIRE(h_scratch != nullptr, kErrBadArg, "regressor: h scratch [B,P] required");
    const size_t smem1 = sizeof(float) * (size_t)HS * F;
    if (smem1 > 48 * 1024) SCAT_ENSURE_SMEM(regressor_hoist_kernel, smem1);
    SCAT_CHECK_CUDA(launch_k(regressor_hoist_kernel, dim3(dim3(ceil_div(B, HS), ceil_div(P, HJ))), dim3(256), smem1, stream, main_feat, Wr, br, h_scratch, B, F, P));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_regressor_fwd(const float* main_feat, const float* feat_out, const float* mean_params, const float* Wr,
                         const float* br, float* pred, float* states, float* h_scratch, int B, int F, int P,
                         int iteration, int root_relative, cudaStream_t stream, int hoisted) {
    SCAT_REQUIRE(P >= 4 && P <= MAXP && F >= 2 && F <= 4096 && (F & 1) == 0, kErrUnsupported,
                 "regressor: P=%d F=%d out of range (F even, P<=96, F<=4096)", P, F);
    SCAT_REQUIRE(!root_relative || (P - 3) % 3 == 0, kErrBadArg, "regressor: root_relative needs P=3+3k");
    SCAT_REQUIRE(h_scratch != nullptr, kErrBadArg, "regressor: h scratch [B,P] required");
    if (!hoisted) SCAT_PROPAGATE(launch_regressor_hoist(main_feat, Wr, br, h_scratch, B, F, P, stream));
    const size_t smem2 = sizeof(float) * ((size_t)P * (P + 1) + 3 * (size_t)P);
    SCAT_CHECK_CUDA(launch_k(regressor_iter_kernel, dim3(B), dim3(128), smem2, stream, h_scratch, feat_out, mean_params, Wr, pred, states, F, P, iteration,
                                                    root_relative));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_regressor_bwd(const float* g_pred, const float* Wr, float* d_feat_out, float* d_main_feat, float* gsum,
                         float* gsteps, int B, int F, int P, int iteration, int root_relative, cudaStream_t stream,
                         int skip_main_feat_gemm) {
    SCAT_REQUIRE(P >= 4 && P <= MAXP, kErrUnsupported, "regressor bwd: P=%d out of range", P);
    SCAT_REQUIRE(gsum && gsteps, kErrBadArg, "regressor bwd: gsum/gsteps scratch required");
    const size_t smem = sizeof(float) * ((size_t)P * (P + 1) + 2 * (size_t)P + 4);
    SCAT_CHECK_CUDA(launch_k(regressor_iter_bwd_kernel, dim3(B), dim3(128), smem, stream, g_pred, Wr, d_feat_out, gsum, gsteps, F, P, iteration, root_relative));
    SCAT_CHECK_LAUNCH();
    if (d_main_feat != nullptr && !skip_main_feat_gemm) {
        GemmArgs g;   // d main_feat[B,F] = gsum[B,P] Wr[:, :F]
        g.A = gsum; g.sam = P; g.sak = 1; g.B = Wr; g.sbn = 1; g.sbk = F + P;
        g.C = d_main_feat; g.ldc = F; g.M = B; g.N = F; g.K = P;
        SCAT_PROPAGATE(launch_gemm_simt(g, stream));
    }
    return 0;
}

}  // namespace scat
