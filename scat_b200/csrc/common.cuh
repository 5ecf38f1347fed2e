// Shared device/host helpers for the scat_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace scat {

// ---- error plumbing: every C-ABI entry returns 0 or a negative/`cudaError_t` code, never throws ----
void set_last_error(const char* fmt, ...);
const char* last_error();

#define SCAT_CHECK_CUDA(expr)                                                                  \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) {                                                               \
            scat::set_last_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,                 \
                                 cudaGetErrorString(_e));                                      \
            return (int)_e;                                                                    \
        }                                                                                      \
    } while (0)

// every kernel launch goes through this macro, so the counter is an exact count of launched kernels
extern unsigned long long g_launch_count;
#define SCAT_CHECK_LAUNCH()                      \
    do {                                         \
        ++scat::g_launch_count;                  \
        SCAT_CHECK_CUDA(cudaGetLastError());     \
    } while (0)

#define SCAT_REQUIRE(cond, code, ...)                                                          \
    do {                                                                                       \
        if (!(cond)) {                                                                         \
            scat::set_last_error(__VA_ARGS__);                                                 \
            return (code);                                                                     \
        }                                                                                      \
    } while (0)

#define SCAT_PROPAGATE(expr)                                                                   \
    do {                                                                                       \
        int _rc = (expr);                                                                      \
        if (_rc != 0) return _rc;                                                              \
    } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (kernel, device): the attribute is per device, and a
// process may drive several GPUs from several host threads (head.cu)
int ensure_dynamic_smem(const void* kernel, int bytes);
extern int g_carveout;
void ensure_carveout(const void* kernel, int pct = 100);
#define SCAT_ENSURE_SMEM(kernel, bytes) SCAT_PROPAGATE(scat::ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), (int)(bytes)))

constexpr int kErrBadArg = -1;
constexpr int kErrWorkspace = -2;
constexpr int kErrUnsupported = -3;

// ---- launches: every kernel goes out with the programmatic-dependent-launch attribute (SCAT_PDL=0 disables) so
// that the next kernel's launch latency and prologue overlap the tail of the current one.  Every kernel calls
// pdl_sync() before it touches global memory: it waits for the preceding grid to complete (and flush), then lets
// the following grid start launching.  Without the attribute both instructions are no-ops.
extern int g_use_pdl;
// Launch priorities: the step is a latency-bound critical chain on the caller's stream plus independent work (weight
// gradients, column sums, weight copies, loss values) on the library's side stream.  Both compete for the same SMs, so
// kernels that go to the side stream are launched with the LOWEST priority and everything else with the highest: the
// block scheduler then serves the critical chain first (the attribute is recorded into CUDA-graph kernel nodes too).
// tl_side_stream is the side stream of the whole-head call in progress on this host thread (null outside one).
extern thread_local cudaStream_t tl_side_stream;
extern int g_prio_low, g_prio_high;     // cudaDeviceGetStreamPriorityRange; equal when SCAT_PRIORITIES=0
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int n = 0;
    if (g_use_pdl) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (g_prio_low != g_prio_high) {
        attr[n].id = cudaLaunchAttributePriority;
        attr[n].val.priority = (tl_side_stream != nullptr && stream == tl_side_stream) ? g_prio_low : g_prio_high;
        ++n;
    }
    if (g_carveout) ensure_carveout(reinterpret_cast<const void*>(kernel));
    cfg.attrs = attr;
    cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// How a producer kernel stores a tensor that feeds a tensor-core GEMM (`out_mode` arguments):
//   OUT_F32  plain fp32;  OUT_TF32  fp32 rounded to the nearest TF32 value (the tensor core would truncate);
//   OUT_BF16 bf16 written through the same pointer (the buffer is then a bf16 array with the given leading
//   dimension in ELEMENTS; it occupies the first half of its fp32-sized workspace slot).
enum OutMode : int { OUT_F32 = 0, OUT_TF32 = 1, OUT_BF16 = 2 };

static inline int ceil_div(int a, int b) { return (a + b - 1) / b; }
static inline size_t round_up(size_t a, size_t b) { return (a + b - 1) / b * b; }

#ifdef __CUDACC__
// packed fp32 FMA (FFMA2, sm_100): (d0, d1) += w * (b0, b1); each lane is an ordinary fma.rn, one issue slot for two
__device__ __forceinline__ void ffma2(float& d0, float& d1, float w, float b0, float b1) {
    unsigned long long d, a, b;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(a) : "f"(w));
    asm("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(b0), "f"(b1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(a), "l"(b));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ float round_tf32_dev(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
// element `idx` of an output tensor stored in `mode`
__device__ __forceinline__ void store_out(float* base, long long idx, float v, int mode) {
    if (mode == OUT_BF16) reinterpret_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
    else base[idx] = mode == OUT_TF32 ? round_tf32_dev(v) : v;
}
// four consecutive elements starting at `idx` (idx % 4 == 0, base 16-byte aligned)
__device__ __forceinline__ void store_out4(float* base, long long idx, float4 v, int mode) {
    if (mode == OUT_BF16) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y), hi = __floats2bfloat162_rn(v.z, v.w);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
        *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(base) + idx) = pk;
    } else {
        if (mode == OUT_TF32) { v.x = round_tf32_dev(v.x); v.y = round_tf32_dev(v.y); v.z = round_tf32_dev(v.z); v.w = round_tf32_dev(v.w); }
        *reinterpret_cast<float4*>(base + idx) = v;
    }
}
// exact (erf) GELU, the nn.GELU() default used by vision_transformer.py:33
__device__ __forceinline__ float gelu_erf(float x) {
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
// gelu(x) and gelu'(x) from one erf: the same expressions as gelu_erf / gelu_erf_grad (0.5 x (1 + e) == x (0.5 (1 + e)) exactly)
__device__ __forceinline__ float gelu_erf_both(float x, float& grad) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    grad = cdf + x * pdf;
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
    const float cdf = 0.5f * (1.0f + erff(x * 0.70710678118654752440f));
    const float pdf = 0.39894228040143267794f * __expf(-0.5f * x * x);
    return cdf + x * pdf;
}
#endif

}  // namespace scat
