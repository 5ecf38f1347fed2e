// fp32-grade GEMM on the legacy tensor path: C = epilogue(A * B^T) with mma.sync.m16n8k8 TF32 and the 3xTF32
// split (a = a_hi + a_lo, b = b_hi + b_lo, a*b ~ a_hi*b_hi + a_lo*b_hi + a_hi*b_lo, fp32 accumulation), whose
// error (~2^-21 relative) is that of fp32 FFMA.
//
// MEASURED AND NOT USED BY THE HEAD: on B200 the legacy mma.sync path issues one m16n8k8 TF32 instruction per ~55
// cycles per SM sub-partition (FFMA-class throughput), and the split needs three per product, so this kernel is
// slower than the fp32 FFMA kernel it was written to replace (44 us vs 25 us on 2016x147x196).  It stays reachable as
// scat_gemm(precision = SCAT_PREC_TF32X3) with its parity tests: arbitrary strides, no alignment requirement.
//
// Motivation: the head keeps three small contractions in fp32 in every precision because they set the output error
// (SURVEY.md section 7): the last feed-forward (196 -> 147 -> 3, vision_transformer.py:37-42), its backward, and
// the regressor weight gradients.  They are far below a tcgen05 tile grid (N = 3, N = 147, K = 3 ...), and as FFMA
// they cost 12 % of the train step.  Here one warp owns a 32 x 32 output tile and loads its operand fragments
// straight from global memory (everything involved is L1/L2 resident: <= 3 MB), with fully general strides
// (A(m,k) = A[m*sam + k*sak]), so forward / dgrad / wgrad, padded leading dimensions and N = 3 all take the same
// path with no staging, no alignment requirement and no block-level synchronisation.  Long-K weight gradients
// split K over gridDim.z and combine with fp32 atomics.
#include "kernels.h"

namespace scat {
namespace {

constexpr int WM = 32, WN = 32;          // warp tile
constexpr int CTA_WARPS = 4;             // 2 x 2 warps: 64 x 64 CTA tile (L1 locality only, the warps never synchronise)

__device__ __forceinline__ uint32_t f2tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return u;
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = f2tf32(x);
    lo = f2tf32(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma8(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

struct RawFrags { float a[2][4]; float b[4][2]; };

// raw fp32 fragment values of one k-step (8 wide) of this warp's 32 x 32 tile; out-of-range elements are zero.
// Every load is unconditional on a clamped (always valid) address and masked afterwards: a predicated load becomes
// a branch, and the compiler then serialises the 16 loads of a k-step behind their uses (measured: 6 us per group).
__device__ __forceinline__ void load_frags(RawFrags& f, const float* __restrict__ A, long long sam, long long sak,
                                           const float* __restrict__ B, long long sbn, long long sbk, int m0, int n0, int k0,
                                           int M, int N, int K, int g, int t) {
    const bool k_lo = k0 + t < K, k_hi = k0 + t + 4 < K;
    const int kc_lo = max(min(k0 + t, K - 1), 0), kc_hi = max(min(k0 + t + 4, K - 1), 0);
    const long long ka_lo = (long long)kc_lo * sak, ka_hi = (long long)kc_hi * sak;
    const long long kb_lo = (long long)kc_lo * sbk, kb_hi = (long long)kc_hi * sbk;
    float va[2][4], vb[4][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const int r0 = min(m0 + mt * 16 + g, M - 1), r1 = min(m0 + mt * 16 + g + 8, M - 1);
        va[mt][0] = __ldg(A + (long long)r0 * sam + ka_lo);
        va[mt][1] = __ldg(A + (long long)r1 * sam + ka_lo);
        va[mt][2] = __ldg(A + (long long)r0 * sam + ka_hi);
        va[mt][3] = __ldg(A + (long long)r1 * sam + ka_hi);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int c = min(n0 + nt * 8 + g, N - 1);
        vb[nt][0] = __ldg(B + (long long)c * sbn + kb_lo);
        vb[nt][1] = __ldg(B + (long long)c * sbn + kb_hi);
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt) {
        const bool r0 = m0 + mt * 16 + g < M, r1 = m0 + mt * 16 + g + 8 < M;
        f.a[mt][0] = (r0 && k_lo) ? va[mt][0] : 0.f;
        f.a[mt][1] = (r1 && k_lo) ? va[mt][1] : 0.f;
        f.a[mt][2] = (r0 && k_hi) ? va[mt][2] : 0.f;
        f.a[mt][3] = (r1 && k_hi) ? va[mt][3] : 0.f;
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const bool c = n0 + nt * 8 + g < N;
        f.b[nt][0] = (c && k_lo) ? vb[nt][0] : 0.f;
        f.b[nt][1] = (c && k_hi) ? vb[nt][1] : 0.f;
    }
}

__global__ void __launch_bounds__(CTA_WARPS * 32)
gemm_mma3_kernel(const GemmArgs g, int ksteps_per_split) {
    pdl_sync();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t = lane & 3;
    const int m0 = blockIdx.y * (2 * WM) + (warp >> 1) * WM;
    const int n0 = blockIdx.x * (2 * WN) + (warp & 1) * WN;
    if (m0 >= g.M || n0 >= g.N) return;
    const float* A = static_cast<const float*>(g.A);
    const float* B = static_cast<const float*>(g.B);
    const int ksteps = (g.K + 7) / 8;
    const int ks_beg = blockIdx.z * ksteps_per_split, ks_end = min(ksteps, ks_beg + ksteps_per_split);
    if (ks_beg >= ks_end) return;

    // Two accumulator sets.  The tensor core adds into its fp32 accumulator with truncation, a bias that grows linearly
    // with the number of chained MMAs (1.5e-5 relative after K = 2016), so a chunk accumulator collects GROUP k-steps
    // (12 MMAs per tile) and is then added to the running total by an ordinary round-to-nearest FADD.
    // The loads of a whole group (64 per lane) are issued before its first MMA: one exposed L2 round trip per group.
    constexpr int GROUP = 4;
    float tot[2][4][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) tot[mt][nt][0] = tot[mt][nt][1] = tot[mt][nt][2] = tot[mt][nt][3] = 0.f;

    for (int ks0 = ks_beg; ks0 < ks_end; ks0 += GROUP) {
        RawFrags f[GROUP];
#pragma unroll
        for (int u = 0; u < GROUP; ++u)      // k-steps past the end read k >= K... only if inside this split's range
            load_frags(f[u], A, g.sam, g.sak, B, g.sbn, g.sbk, m0, n0, (ks0 + u) * 8, g.M, g.N, (ks0 + u < ks_end) ? g.K : 1, gq, t);
        float acc[2][4][4];
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) acc[mt][nt][0] = acc[mt][nt][1] = acc[mt][nt][2] = acc[mt][nt][3] = 0.f;
#pragma unroll
        for (int u = 0; u < GROUP; ++u) {
            uint32_t ah[2][4], al[2][4], bh[4][2], bl[4][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int e = 0; e < 4; ++e) split_tf32(f[u].a[mt][e], ah[mt][e], al[mt][e]);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 2; ++e) split_tf32(f[u].b[nt][e], bh[nt][e], bl[nt][e]);
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    mma8(acc[mt][nt], al[mt], bh[nt]);      // small terms first
                    mma8(acc[mt][nt], ah[mt], bl[nt]);
                    mma8(acc[mt][nt], ah[mt], bh[nt]);
                }
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) tot[mt][nt][e] += acc[mt][nt][e];
    }

    const bool split = gridDim.z > 1;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int m = m0 + mt * 16 + gq + (e >> 1) * 8;
                const int n = n0 + nt * 8 + 2 * t + (e & 1);
                if (m >= g.M || n >= g.N) continue;
                float v = tot[mt][nt][e];
                const long long ma = g.aux_row_mod > 0 ? m % g.aux_row_mod : m;
                switch (g.epilogue) {
                    case EPI_BIAS: v += g.bias[n]; break;
                    case EPI_BIAS_RESID: v += g.bias[n] + g.aux_in[ma * g.ld_aux_in + n]; break;
                    case EPI_BIAS_GELU: {
                        v += g.bias[n];
                        g.aux_out[(long long)m * g.ld_aux_out + n] = v;
                        v = gelu_erf(v);
                    } break;
                    case EPI_DGELU: v *= gelu_erf_grad(g.aux_in[ma * g.ld_aux_in + n]); break;
                    case EPI_RESID: v += g.aux_in[ma * g.ld_aux_in + n]; break;
                    default: break;
                }
                if (g.round_out) v = round_tf32(v);
                float* c = g.C + (long long)m * g.ldc + n;
                if (split) atomicAdd(c, v);                       // C was cleared (or holds the accumulate base)
                else *c = g.accumulate ? (*c + v) : v;
                if (g.C16 != nullptr) reinterpret_cast<__nv_bfloat16*>(g.C16)[(long long)m * g.ldc16 + n] = __float2bfloat16_rn(v);
            }
}

}  // namespace

int launch_gemm_mma3(const GemmArgs& g, cudaStream_t stream) {
    SCAT_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, kErrBadArg, "gemm: bad shape %d %d %d", g.M, g.N, g.K);
    SCAT_REQUIRE(g.A && g.B && g.C, kErrBadArg, "gemm: null operand");
    SCAT_REQUIRE(!g.operand_bf16, kErrUnsupported, "gemm_mma3 takes fp32 operands");
    if (g.epilogue == EPI_BIAS || g.epilogue == EPI_BIAS_RESID || g.epilogue == EPI_BIAS_GELU)
        SCAT_REQUIRE(g.bias != nullptr, kErrBadArg, "gemm: epilogue %d needs bias", g.epilogue);
    if (g.epilogue == EPI_BIAS_RESID || g.epilogue == EPI_DGELU || g.epilogue == EPI_RESID)
        SCAT_REQUIRE(g.aux_in != nullptr, kErrBadArg, "gemm: epilogue %d needs aux_in", g.epilogue);
    if (g.epilogue == EPI_BIAS_GELU)
        SCAT_REQUIRE(g.aux_out != nullptr, kErrBadArg, "gemm: epilogue %d needs aux_out", g.epilogue);
    SCAT_REQUIRE(g.epilogue <= EPI_RESID, kErrUnsupported, "gemm_mma3: epilogue %d", g.epilogue);
    dim3 grid(ceil_div(g.N, 2 * WN), ceil_div(g.M, 2 * WM), 1);
    const int ksteps = ceil_div(g.K, 8);
    int splits = 1;
    const int tiles = (int)(grid.x * grid.y);
    if (g.allow_split_k && g.epilogue == EPI_NONE && !g.C16 && !g.round_out && tiles < 148 && ksteps >= 16)
        splits = max(1, min(ksteps / 8, ceil_div(296, tiles)));
    const int per = ceil_div(ksteps, splits);
    grid.z = ceil_div(ksteps, per);
    if (grid.z > 1 && !g.accumulate && !g.c_zeroed)
        SCAT_CHECK_CUDA(cudaMemset2DAsync(g.C, (size_t)g.ldc * sizeof(float), 0, (size_t)g.N * sizeof(float), g.M, stream));
    SCAT_CHECK_CUDA(launch_k(gemm_mma3_kernel, dim3(grid), dim3(CTA_WARPS * 32), 0, stream, g, per));
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace scat
