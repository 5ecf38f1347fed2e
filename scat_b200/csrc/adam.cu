// Fused Adam step for the head's parameters (replaces optim.Adam(...).step(),
// /root/reference/train.py:60,209, for the head's 35 tensors -- the backbone keeps its PyTorch optimiser).
//
// Parameters, gradients and both moments are flat fp32 buffers in the same order (the gradient bucket's order), so
// the whole update is ONE elementwise pass: 16 B/element read (p, g, m, v) + 12 B/element written (p, m, v) =
// 28 B/element, 106 MB for the 3,795,099-element head -- HBM bound.  The arithmetic follows torch's single-tensor
// Adam operation by operation (oracle/adam_oracle.py); the scalar bias corrections are computed in double once per
// block.  The step count and learning rate can be read from device memory so that a captured CUDA graph replays with
// a changing schedule.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace scat {
namespace {

struct AdamScalars {
    float one_minus_b1, b2, one_minus_b2, neg_step_size, bc2_sqrt, eps, wd;
};

__device__ __forceinline__ void adam_elem(float& p, float g, float& m, float& v, const AdamScalars& s) {
    if (s.wd != 0.f) g = fmaf(s.wd, p, g);
    m = fmaf(s.one_minus_b1, g - m, m);
    v = fmaf(s.one_minus_b2 * g, g, v * s.b2);
    const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), s.bc2_sqrt), s.eps);
    p = __fadd_rn(p, __fdiv_rn(__fmul_rn(s.neg_step_size, m), denom));      // addcdiv_: self + value * t1 / t2
}

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                   float* __restrict__ v, long long n, double lr, double beta1, double beta2,
                                                   float eps, float wd, int step, const float* __restrict__ lr_dev,
                                                   const int* __restrict__ step_dev, const uint32_t* __restrict__ abort_flag) {
    pdl_sync();
    __shared__ AdamScalars sh;
    __shared__ uint32_t aborted;
    if (threadIdx.x == 0) aborted = abort_flag ? *reinterpret_cast<const volatile uint32_t*>(abort_flag) : 0u;
    __syncthreads();
    if (aborted) return;          // the gradient exchange before this update gave up on a peer: leave the weights alone
    if (threadIdx.x == 0) {
        const double t = (double)(step_dev ? *step_dev : step);
        // betas arrive in double: 1 - (float)0.999 differs from (float)(1 - 0.999) by 1.3e-5 relative
        const double rate = lr_dev ? (double)*lr_dev : lr;
        const double bc1 = 1.0 - pow(beta1, t), bc2 = 1.0 - pow(beta2, t);
        sh.one_minus_b1 = (float)(1.0 - beta1);
        sh.b2 = (float)beta2;
        sh.one_minus_b2 = (float)(1.0 - beta2);
        sh.neg_step_size = (float)(-rate / bc1);
        sh.bc2_sqrt = (float)sqrt(bc2);
        sh.eps = eps;
        sh.wd = wd;
    }
    __syncthreads();
    const AdamScalars s = sh;
    const long long n4 = n / 4, stride = (long long)gridDim.x * blockDim.x;
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long i = tid; i < n4; i += stride) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        adam_elem(pp.x, gg.x, mm.x, vv.x, s);
        adam_elem(pp.y, gg.y, mm.y, vv.y, s);
        adam_elem(pp.z, gg.z, mm.z, vv.z, s);
        adam_elem(pp.w, gg.w, mm.w, vv.w, s);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
    }
    for (long long i = n4 * 4 + tid; i < n; i += stride) adam_elem(p[i], g[i], m[i], v[i], s);
}

}  // namespace
}  // namespace scat

using namespace scat;

extern "C" int scat_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, long long n,
                              double lr, double beta1, double beta2, double eps, double weight_decay, int32_t step,
                              const float* lr_dev, const int32_t* step_dev, const uint32_t* abort_flag, void* stream) {
    SCAT_REQUIRE(params && grads && exp_avg && exp_avg_sq && n > 0, kErrBadArg, "adam_step: null buffer or n <= 0");
    SCAT_REQUIRE(step_dev || step >= 1, kErrBadArg, "adam_step: step %d (1-based count of this update)", step);
    SCAT_REQUIRE(beta1 >= 0.0 && beta1 < 1.0 && beta2 >= 0.0 && beta2 < 1.0 && eps >= 0.0, kErrBadArg,
                 "adam_step: betas (%g, %g) eps %g", beta1, beta2, eps);
    const uintptr_t al = (uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq;
    SCAT_REQUIRE(al % 16 == 0, kErrBadArg, "adam_step: buffers must be 16-byte aligned");
    // exactly one resident wave (grid-stride loop): ncu showed 1.6 waves, i.e. a 0.6-wave tail, with a fixed 8 CTAs per SM
    static int resident_of[16] = {0};                  // per device (a process may drive several GPUs)
    int dev = 0;
    SCAT_CHECK_CUDA(cudaGetDevice(&dev));
    int resident = (dev >= 0 && dev < 16) ? resident_of[dev] : 0;
    if (resident == 0) {
        int sms = 0, per_sm = 0;
        SCAT_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        SCAT_CHECK_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, adam_kernel, 256, 0));
        resident = std::max(1, sms * per_sm);
        if (dev >= 0 && dev < 16) resident_of[dev] = resident;
    }
    const long long n4 = (n + 3) / 4;
    const int grid = (int)std::min<long long>(resident, (n4 + 255) / 256);
    SCAT_CHECK_CUDA(launch_k(adam_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, params, grads, exp_avg, exp_avg_sq,
                             n, lr, beta1, beta2, (float)eps, (float)weight_decay, (int)step, lr_dev, (const int*)step_dev, abort_flag));
    SCAT_CHECK_LAUNCH();
    return 0;
}
