// Evaluation metrics on the device: the step after the head at inference.  The reference
// moves every frame to numpy for these (eval.py:691-753); here the joints stay on the GPU.
//   procrustes_kernel    batch_compute_similarity_transform_torch   /root/reference/eval.py:110-161
//   joint_error_kernel   distances + threshold counts of cal_PCK     /root/reference/eval.py:300-316, MPJPE eval.py:749
//   accel_kernel         compute_accel / compute_error_accel         /root/reference/data_utils/eval_utils.py:6-47
// Tiny, latency-bound kernels (63 floats per sample); the arithmetic lives in eval_math.cuh, shared with the host test.
#include "common.cuh"
#include "eval_math.cuh"
#include "kernels.h"

namespace scat {
namespace {

constexpr int kProcThreads = 64;
constexpr int kMaxThresholds = 64;

struct Thresholds {
    double v[kMaxThresholds];
    int n;
};

// one thread per sample; the block's samples are staged through shared memory so global accesses are coalesced
// (row stride 3n is odd for n = 21: conflict-free per-thread rows)
__global__ void __launch_bounds__(kProcThreads) procrustes_kernel(const float* __restrict__ pred, const float* __restrict__ gt,
                                                                 int batch, int n, float* __restrict__ aligned,
                                                                 float* __restrict__ scale) {
    pdl_sync();
    extern __shared__ float sm[];
    const int row = 3 * n, b0 = blockIdx.x * kProcThreads;
    const int rows = min(kProcThreads, batch - b0);
    float* s1 = sm;
    float* s2 = sm + kProcThreads * row;
    for (int i = threadIdx.x; i < rows * row; i += kProcThreads) {
        s1[i] = pred[(size_t)b0 * row + i];
        s2[i] = gt[(size_t)b0 * row + i];
    }
    __syncthreads();
    if (threadIdx.x < rows) {
        float sc;
        evalm::similarity_align(s1 + threadIdx.x * row, s2 + threadIdx.x * row, n, s1 + threadIdx.x * row, &sc);
        if (scale) scale[b0 + threadIdx.x] = sc;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < rows * row; i += kProcThreads) aligned[(size_t)b0 * row + i] = s1[i];
}

// one thread per (sample, joint), whole samples per block (S = 256 / n of them): d = sqrt(sum_c (unit*p - unit*g)^2)
// with torch's rounding points (separate multiplies, subtract, squares added left to right); counts[k] += d <=
// thresholds[k] (compared in double like numpy) through per-block shared counters; mpjpe[b] = the per-sample mean of the
// un-scaled distance, summed in joint order by one thread (deterministic)
__global__ void __launch_bounds__(256) joint_error_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int batch,
                                                          int n, int S, float unit, Thresholds thr,
                                                          unsigned long long* __restrict__ counts, float* __restrict__ mpjpe) {
    pdl_sync();
    __shared__ float sraw[256];
    __shared__ unsigned int sc[kMaxThresholds];
    if (threadIdx.x < kMaxThresholds) sc[threadIdx.x] = 0u;
    __syncthreads();
    const int ls = threadIdx.x / n, j = threadIdx.x - ls * n;
    const long long b = (long long)blockIdx.x * S + ls;
    const bool live = ls < S && b < batch;
    float d = 0.f, raw = 0.f;
    if (live) {
        const long long i = b * n + j;
        float acc = 0.f, acc_raw = 0.f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float p = pred[i * 3 + c], g = gt[i * 3 + c];
            const float e = __fsub_rn(__fmul_rn(p, unit), __fmul_rn(g, unit));
            acc = __fadd_rn(acc, __fmul_rn(e, e));
            const float r = __fsub_rn(p, g);
            acc_raw = __fadd_rn(acc_raw, __fmul_rn(r, r));
        }
        d = __fsqrt_rn(acc);
        raw = __fsqrt_rn(acc_raw);
    }
    sraw[threadIdx.x] = raw;
    for (int k = 0; k < thr.n; ++k) {
        const unsigned hit = __ballot_sync(0xffffffffu, live && (double)d <= thr.v[k]);
        if ((threadIdx.x & 31) == 0 && hit) atomicAdd(&sc[k], (unsigned)__popc(hit));
    }
    __syncthreads();
    if (mpjpe && live && j == 0) {
        float sum = 0.f;
        for (int jj = 0; jj < n; ++jj) sum = __fadd_rn(sum, sraw[threadIdx.x + jj]);
        mpjpe[b] = sum / (float)n;
    }
    if (threadIdx.x < thr.n && sc[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)sc[threadIdx.x]);
}

// one thread per (frame triple, joint): || (p[i] - 2 p[i+1] + p[i+2]) - (g[i] - 2 g[i+1] + g[i+2]) ||, averaged over
// the joints (gt == nullptr: the plain acceleration of compute_accel)
__global__ void __launch_bounds__(256) accel_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int frames, int n,
                                                    float* __restrict__ out) {
    pdl_sync();
    const long long total = (long long)(frames - 2) * n;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long f = i / n, j = i % n, stride = (long long)n * 3;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const long long o = f * stride + j * 3 + c;
        float a = __fadd_rn(__fsub_rn(pred[o], __fmul_rn(2.f, pred[o + stride])), pred[o + 2 * stride]);
        if (gt) a = __fsub_rn(a, __fadd_rn(__fsub_rn(gt[o], __fmul_rn(2.f, gt[o + stride])), gt[o + 2 * stride]));
        acc = __fadd_rn(acc, __fmul_rn(a, a));
    }
    atomicAdd(&out[f], __fsqrt_rn(acc) / (float)n);
}

}  // namespace
}  // namespace scat

using namespace scat;

extern "C" {

int scat_eval_procrustes(const float* pred, const float* gt, int32_t batch, int32_t n_joints, float* aligned, float* scale,
                         void* stream) {
    SCAT_REQUIRE(pred && gt && aligned, kErrBadArg, "eval_procrustes: null tensor");
    SCAT_REQUIRE(batch > 0 && n_joints >= 3 && n_joints <= 64, kErrBadArg, "eval_procrustes: batch %d joints %d (3..64)", batch,
                 n_joints);
    const size_t smem = (size_t)2 * kProcThreads * 3 * n_joints * sizeof(float);
    SCAT_ENSURE_SMEM(procrustes_kernel, 2 * kProcThreads * 3 * 64 * 4);
    SCAT_CHECK_CUDA(launch_k(procrustes_kernel, dim3(ceil_div(batch, kProcThreads)), dim3(kProcThreads), smem, (cudaStream_t)stream,
                             pred, gt, (int)batch, (int)n_joints, aligned, scale));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int scat_eval_joint_errors(const float* pred, const float* gt, int32_t batch, int32_t n_joints, float unit_scale,
                           const double* thresholds, int32_t n_thresholds, unsigned long long* counts, float* mpjpe,
                           void* stream) {
    SCAT_REQUIRE(pred && gt, kErrBadArg, "eval_joint_errors: null tensor");
    SCAT_REQUIRE(batch > 0 && n_joints > 0 && n_joints <= 256, kErrBadArg, "eval_joint_errors: batch %d joints %d (1..256)", batch,
                 n_joints);
    SCAT_REQUIRE(n_thresholds >= 0 && n_thresholds <= kMaxThresholds && (n_thresholds == 0 || (thresholds && counts)), kErrBadArg,
                 "eval_joint_errors: %d thresholds (at most %d, with a counts buffer)", n_thresholds, kMaxThresholds);
    cudaStream_t st = (cudaStream_t)stream;
    Thresholds thr = {};
    thr.n = n_thresholds;
    for (int k = 0; k < n_thresholds; ++k) thr.v[k] = thresholds[k];
    if (n_thresholds) SCAT_CHECK_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * n_thresholds, st));
    const int S = 256 / n_joints;                       // whole samples per block
    SCAT_CHECK_CUDA(launch_k(joint_error_kernel, dim3((unsigned)ceil_div(batch, S)), dim3(256), 0, st, pred, gt, (int)batch,
                             (int)n_joints, S, unit_scale, thr, counts, mpjpe));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int scat_eval_accel(const float* pred, const float* gt, int32_t n_frames, int32_t n_joints, float* out, void* stream) {
    SCAT_REQUIRE(pred && out, kErrBadArg, "eval_accel: null tensor");
    SCAT_REQUIRE(n_frames >= 3 && n_joints > 0, kErrBadArg, "eval_accel: %d frames (at least 3), %d joints", n_frames, n_joints);
    cudaStream_t st = (cudaStream_t)stream;
    SCAT_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (n_frames - 2), st));
    const long long total = (long long)(n_frames - 2) * n_joints;
    SCAT_CHECK_CUDA(launch_k(accel_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, pred, gt, (int)n_frames,
                             (int)n_joints, out));
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // extern "C"
