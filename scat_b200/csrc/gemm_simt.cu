// fp32 CUDA-core GEMM with strided operands and fused epilogues ("parity mode", and the path for
// shapes the tensor-core kernel does not take: N=3, N=147/K=147 last feed-forward, regressor grads).
//
// Replaces the cuBLAS calls behind nn.Linear on the reference's path (vision_transformer.py:33-35,53-55)
// and their autograd backward.  C = epilogue(A * B^T) with A(m,k) and B(n,k) addressed through
// (row, col) strides so forward / dgrad / wgrad are the same kernel (see kernels.h).
// Small-output / long-K problems (weight gradients: K = B*21 rows) are split along K over gridDim.z and
// combined with fp32 atomics so that they still fill the 148 SMs.
#include "kernels.h"

namespace scat {

namespace {

constexpr int BM = 64, BN = 64, BK = 16, THREADS = 256, PAD = 4;

constexpr int STAGES = 4;     // cp.async ring: the K loop of these small GEMMs is latency bound, not FMA bound

__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src, bool valid) {
    const unsigned bytes = valid ? 4u : 0u;       // 0 source bytes = zero fill
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src, bool valid) {
    const unsigned bytes = valid ? 16u : 0u;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// Stage a [64 rows x 16 k] operand tile into S[k][row] with asynchronous copies.  Two thread->element maps so that
// the unit-stride direction is the fastest-varying one across a warp (coalesced either way); the K-contiguous map
// transposes on the fly (4-byte copies), the row-contiguous one moves 16 bytes when the operand allows it.
template <bool ROW_CONTIG_K>
__device__ __forceinline__ void stage_tile(const float* __restrict__ P, long long s_row, long long s_k, int row0, int k0,
                                           int rows, int K, int tid, float (*S)[BM + PAD], bool vec16) {
    if (ROW_CONTIG_K) {
        const int r = tid >> 2, kk = (tid & 3) * 4;
        const int gr = row0 + r, gk = k0 + kk;
        const float* p = P + (long long)gr * s_row + (long long)gk * s_k;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const bool ok = gr < rows && gk + i < K;
            cp_async4(&S[kk + i][r], ok ? p + i * s_k : P, ok);
        }
    } else {
        const int kk = tid >> 4, r = (tid & 15) * 4;
        const int gk = k0 + kk, gr = row0 + r;
        const float* p = P + (long long)gr * s_row + (long long)gk * s_k;
        if (vec16 && gk < K && gr + 3 < rows) {
            cp_async16(&S[kk][r], p, true);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const bool ok = gk < K && gr + i < rows;
                cp_async4(&S[kk][r + i], ok ? p + i * s_row : P, ok);
            }
        }
    }
}

template <bool A_K, bool B_K>
__global__ void __launch_bounds__(THREADS) gemm_simt_kernel(GemmArgs g, int kt_per_split) {
    pdl_sync();
    __shared__ __align__(16) float As[STAGES][BK][BM + PAD];
    __shared__ __align__(16) float Bs[STAGES][BK][BN + PAD];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int ty = tid >> 4, tx = tid & 15;
    const float* A = static_cast<const float*>(g.A);
    const float* B = static_cast<const float*>(g.B);

    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    const int nk_all = (g.K + BK - 1) / BK;
    const int kt_beg = blockIdx.z * kt_per_split;
    const int kt_end = min(nk_all, kt_beg + kt_per_split);
    if (kt_beg >= kt_end) return;
    const int nkt = kt_end - kt_beg;
    // 16-byte copies need a unit row stride, 16-byte aligned k-rows and base
    const bool a16 = !A_K && g.sam == 1 && (g.sak & 3) == 0 && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) && (m0 & 3) == 0;
    const bool b16 = !B_K && g.sbn == 1 && (g.sbk & 3) == 0 && ((reinterpret_cast<uintptr_t>(B) & 15) == 0) && (n0 & 3) == 0;
#pragma unroll
    for (int s = 0; s < STAGES - 1; ++s) {
        if (s < nkt) {
            stage_tile<A_K>(A, g.sam, g.sak, m0, (kt_beg + s) * BK, g.M, g.K, tid, As[s], a16);
            stage_tile<B_K>(B, g.sbn, g.sbk, n0, (kt_beg + s) * BK, g.N, g.K, tid, Bs[s], b16);
        }
        cp_async_commit();
    }
    for (int it = 0; it < nkt; ++it) {
        cp_async_wait<STAGES - 2>();          // the group that filled stage it % STAGES has landed (for this thread)
        __syncthreads();                      // ... and for every thread; everyone is also done computing on stage it-1
        const int nxt = it + STAGES - 1;
        if (nxt < nkt) {
            stage_tile<A_K>(A, g.sam, g.sak, m0, (kt_beg + nxt) * BK, g.M, g.K, tid, As[nxt % STAGES], a16);
            stage_tile<B_K>(B, g.sbn, g.sbk, n0, (kt_beg + nxt) * BK, g.N, g.K, tid, Bs[nxt % STAGES], b16);
        }
        cp_async_commit();
        const int st = it % STAGES;
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[st][k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[st][k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; j += 2) ffma2(acc[i][j], acc[i][j + 1], av[i], bv[j], bv[j + 1]);   // packed FMA: same values
        }
    }

    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            const long long ma = g.aux_row_mod > 0 ? m % g.aux_row_mod : m;
            switch (g.epilogue) {
                case EPI_BIAS: v += g.bias[n]; break;
                case EPI_BIAS_RESID: v += g.bias[n] + g.aux_in[ma * g.ld_aux_in + n]; break;
                case EPI_BIAS_GELU: {
                    v += g.bias[n];
                    if (g.gelu_saves_grad) {
                        float dg;
                        v = gelu_erf_both(v, dg);
                        g.aux_out[(long long)m * g.ld_aux_out + n] = dg;
                    } else {
                        g.aux_out[(long long)m * g.ld_aux_out + n] = v;
                        v = gelu_erf(v);
                    }
                } break;
                case EPI_DGELU: {
                    const float z = g.aux_in[ma * g.ld_aux_in + n];
                    v *= g.gelu_saves_grad ? z : gelu_erf_grad(z);
                } break;
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
}

// column sums out[n] (+)= sum_m X[m,n]: grid (column groups, row slices); slices combine with atomics
__global__ void colsum_kernel(const float* __restrict__ X, int ld, int M, int N, float* __restrict__ out) {
    pdl_sync();
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (c < N)
        for (int m = blockIdx.y * 8 + threadIdx.y; m < M; m += 8 * gridDim.y) s += X[(long long)m * ld + c];
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
        atomicAdd(out + c, t);
    }
}

__global__ void colsum_bf16_kernel(const __nv_bfloat16* __restrict__ X, int ld, int M, int N, float* __restrict__ out) {
    pdl_sync();
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    float s = 0.f;
    if (c < N)
        for (int m = blockIdx.y * 8 + threadIdx.y; m < M; m += 8 * gridDim.y) s += __bfloat162float(X[(long long)m * ld + c]);
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < N) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
        atomicAdd(out + c, t);
    }
}

// grid (x, job): block column x strides job blockIdx.y.  (Measured alternatives, r2v2 / r2x timelines: ONE 256- or 128-thread
// block per SM walking every job, launched so that the conv forward gets the SMs first -- the copies then start 15 us
// after the conv kernel and finish 47 us into the step, delaying the first GEMM; the kernels do not overlap either way.)
__device__ __forceinline__ void round_copy_job(const RoundJob& job, int tid, int nthr);
__global__ void __launch_bounds__(256) round_copy_kernel(const RoundJobs jobs) {
    pdl_sync();
    round_copy_job(jobs.job[blockIdx.y], blockIdx.x * blockDim.x + threadIdx.x, gridDim.x * blockDim.x);
}
__device__ __forceinline__ void round_copy_job(const RoundJob& job, int tid, int nthr) {
    if (job.to_bf16 == 2) {                                          // TF32 remainder: src - round(src), itself rounded
        const int total = job.rows * job.ld_dst;
        for (int i = tid; i < total; i += nthr) {
            const int r = i / job.ld_dst, c = i - r * job.ld_dst;
            const float v = c < job.cols ? __ldg(job.src + (size_t)r * job.ld_src + c) : 0.f;
            job.dst[i] = round_tf32(v - round_tf32(v));
        }
        return;
    }
    if (job.to_bf16 == 1) {                                          // bf16 copy, pad columns zero-filled
        __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(job.dst);
        if (((job.cols | job.ld_src | job.ld_dst) & 3) == 0) {       // 4 values per thread: float4 in, 8 bytes out
            const int c4n = job.ld_dst >> 2, total = job.rows * c4n;
            for (int i0 = tid; i0 < total; i0 += 4 * nthr) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * nthr;
                    const int r = i / c4n, c4 = i - r * c4n;
                    v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (i < total && c4 * 4 < job.cols) v[u] = __ldg(reinterpret_cast<const float4*>(job.src + (size_t)r * job.ld_src) + c4);
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * nthr;
                    if (i < total) {
                        const __nv_bfloat162 lo = __floats2bfloat162_rn(v[u].x, v[u].y), hi = __floats2bfloat162_rn(v[u].z, v[u].w);
                        uint2 pk;
                        pk.x = *reinterpret_cast<const uint32_t*>(&lo);
                        pk.y = *reinterpret_cast<const uint32_t*>(&hi);
                        reinterpret_cast<uint2*>(dst)[i] = pk;
                    }
                }
            }
            return;
        }
        const int total = job.rows * job.ld_dst;
        for (int i = tid; i < total; i += nthr) {
            const int r = i / job.ld_dst, c = i - r * job.ld_dst;
            dst[i] = __float2bfloat16_rn(c < job.cols ? __ldg(job.src + (size_t)r * job.ld_src + c) : 0.f);
        }
        return;
    }
    if (((job.cols | job.ld_src | job.ld_dst) & 3) == 0) {          // float4 path (every weight but fc2 of layer 1)
        // four loads in flight per thread before the first store (dst may alias src as far as the compiler knows, so a
        // one-element loop body waits a DRAM round trip per element: 37 us in the step's timeline for 15 MB)
        const int c4n = job.ld_dst >> 2, total = job.rows * c4n;
        for (int i0 = tid; i0 < total; i0 += 4 * nthr) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * nthr;
                const int r = i / c4n, c4 = i - r * c4n;
                v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (i < total && c4 * 4 < job.cols) v[u] = __ldg(reinterpret_cast<const float4*>(job.src + (size_t)r * job.ld_src) + c4);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u * nthr;
                if (i < total) {
                    float4 w = v[u];
                    w.x = round_tf32(w.x); w.y = round_tf32(w.y); w.z = round_tf32(w.z); w.w = round_tf32(w.w);
                    reinterpret_cast<float4*>(job.dst)[i] = w;
                }
            }
        }
    } else {
        const int total = job.rows * job.ld_dst;
        for (int i = tid; i < total; i += nthr) {
            const int r = i / job.ld_dst, c = i - r * job.ld_dst;
            job.dst[i] = c < job.cols ? round_tf32(__ldg(job.src + (size_t)r * job.ld_src + c)) : 0.f;
        }
    }
}

// fp32-grade tensor-core operands (3xTF32): every value is split into hi = TF32-nearest(v) and lo = TF32-nearest(v - hi),
// both exact on the tensor core, and the two operands of a product are stacked along K so that ONE GEMM computes
// A_hi B_hi + A_lo B_hi + A_hi B_lo (the dropped lo x lo term is 2^-22 relative):
//   first operand  (mode 0): dst[r, 3 x cp] = [hi | lo | hi]           (cp = cols padded to 8, pad columns zero)
//   second operand (mode 1): dst[r, 3 x cp] = [hi | hi | lo]           K-major second operand (forward, y = x W^T)
//   second operand (mode 2): dst[3 x rp, cols] = [hi; hi; lo]          MN-major second operand (dgrad, dx = dy W), rp = rows padded to 8
__global__ void split3_kernel(const float* __restrict__ src, int ld_src, float* __restrict__ dst, int rows, int cols, int mode) {
    pdl_sync();
    const int cp = (cols + 7) & ~7, rp = (rows + 7) & ~7;          // operand pitches: 32-byte rows (head.cu: padp)
    if (mode != 2 && ((cols | ld_src) & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0) {
        // activations (the step's critical chain): one float4 in, three float4 out per thread; pad columns written as zeros
        const int c4n = cp >> 2;
        const long long total4 = (long long)rows * c4n;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (long long)gridDim.x * blockDim.x) {
            const int r = (int)(i / c4n), c4 = (int)(i - (long long)r * c4n);
            const float4 v = c4 * 4 < cols ? __ldg(reinterpret_cast<const float4*>(src + (long long)r * ld_src) + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            float4 hi, lo;
            hi.x = round_tf32(v.x); hi.y = round_tf32(v.y); hi.z = round_tf32(v.z); hi.w = round_tf32(v.w);
            lo.x = round_tf32(v.x - hi.x); lo.y = round_tf32(v.y - hi.y); lo.z = round_tf32(v.z - hi.z); lo.w = round_tf32(v.w - hi.w);
            float4* d = reinterpret_cast<float4*>(dst + (long long)r * 3 * cp) + c4;
            d[0] = hi;
            d[c4n] = mode == 0 ? lo : hi;
            d[2 * c4n] = mode == 0 ? hi : lo;
        }
        return;
    }
    const long long total = mode == 2 ? (long long)rp * cols : (long long)rows * cp;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int width = mode == 2 ? cols : cp;
        const int r = (int)(i / width), c = (int)(i - (long long)r * width);
        const float v = (r < rows && c < cols) ? __ldg(src + (long long)r * ld_src + c) : 0.f;
        const float hi = round_tf32(v), lo = round_tf32(v - hi);
        if (mode == 2) {
            dst[i] = hi;
            dst[(long long)rp * cols + i] = hi;
            dst[2LL * rp * cols + i] = lo;
        } else {
            float* d = dst + (long long)r * 3 * cp + c;
            d[0] = hi;
            d[cp] = mode == 0 ? lo : hi;
            d[2 * cp] = mode == 0 ? hi : lo;
        }
    }
}

}  // namespace

int launch_split3(const float* src, int ld_src, float* dst, int rows, int cols, int mode, cudaStream_t stream) {
    SCAT_REQUIRE(src && dst && rows > 0 && cols > 0 && mode >= 0 && mode <= 2, kErrBadArg, "split3: bad args");
    const long long total = (long long)((rows + 7) & ~7) * ((cols + 7) & ~7);
    const bool vec = mode != 2 && ((cols | ld_src) & 3) == 0;
    const long long work = vec ? total / 4 : total;
    const int grid = (int)((work + 255) / 256 < 148 * 4 ? (work + 255) / 256 : 148 * 4);
    SCAT_CHECK_CUDA(launch_k(split3_kernel, dim3(grid), dim3(256), 0, stream, src, ld_src, dst, rows, cols, mode));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_round_copy(const RoundJobs& jobs, cudaStream_t stream) {
    SCAT_REQUIRE(jobs.n > 0 && jobs.n <= 16, kErrBadArg, "round_copy: %d jobs", jobs.n);
    SCAT_CHECK_CUDA(launch_k(round_copy_kernel, dim3(dim3(74, jobs.n)), dim3(256), 0, stream, jobs));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_gemm_simt(const GemmArgs& g, cudaStream_t stream) {
    SCAT_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, kErrBadArg, "gemm: bad shape %d %d %d", g.M, g.N, g.K);
    SCAT_REQUIRE(g.A && g.B && g.C, kErrBadArg, "gemm: null operand");
    SCAT_REQUIRE(!g.operand_bf16, kErrUnsupported, "gemm: the FFMA kernel takes fp32 operands only");
    if (g.epilogue == EPI_BIAS || g.epilogue == EPI_BIAS_RESID || g.epilogue == EPI_BIAS_GELU)
        SCAT_REQUIRE(g.bias != nullptr, kErrBadArg, "gemm: epilogue %d needs bias", g.epilogue);
    if (g.epilogue == EPI_BIAS_RESID || g.epilogue == EPI_DGELU || g.epilogue == EPI_RESID)
        SCAT_REQUIRE(g.aux_in != nullptr, kErrBadArg, "gemm: epilogue %d needs aux_in", g.epilogue);
    if (g.epilogue == EPI_BIAS_GELU)
        SCAT_REQUIRE(g.aux_out != nullptr, kErrBadArg, "gemm: epilogue %d needs aux_out", g.epilogue);
    dim3 grid(ceil_div(g.N, BN), ceil_div(g.M, BM), 1);
    const int nk = ceil_div(g.K, BK);
    int splits = 1;
    if (g.allow_split_k && g.epilogue == EPI_NONE && !g.C16 && (int)(grid.x * grid.y) < 74 && nk >= 16) {
        splits = min(min(32, nk / 4), ceil_div(296, (int)(grid.x * grid.y)));
        if (splits < 1) splits = 1;
    }
    const int kt_per_split = ceil_div(nk, splits);
    grid.z = ceil_div(nk, kt_per_split);
    if (grid.z > 1 && !g.accumulate && !g.c_zeroed)
        SCAT_CHECK_CUDA(cudaMemset2DAsync(g.C, (size_t)g.ldc * sizeof(float), 0, (size_t)g.N * sizeof(float), g.M, stream));
    const bool a_k = (g.sak == 1) || (g.sam != 1);
    const bool b_k = (g.sbk == 1) || (g.sbn != 1);
    if (a_k && b_k) SCAT_CHECK_CUDA(launch_k(gemm_simt_kernel<true, true>, dim3(grid), dim3(THREADS), 0, stream, g, kt_per_split));
    else if (a_k && !b_k) SCAT_CHECK_CUDA(launch_k(gemm_simt_kernel<true, false>, dim3(grid), dim3(THREADS), 0, stream, g, kt_per_split));
    else if (!a_k && b_k) SCAT_CHECK_CUDA(launch_k(gemm_simt_kernel<false, true>, dim3(grid), dim3(THREADS), 0, stream, g, kt_per_split));
    else SCAT_CHECK_CUDA(launch_k(gemm_simt_kernel<false, false>, dim3(grid), dim3(THREADS), 0, stream, g, kt_per_split));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_colsum(const float* X, int ld, int M, int N, float* out, int accumulate, cudaStream_t stream, int x_bf16) {
    SCAT_REQUIRE(X && out && M > 0 && N > 0, kErrBadArg, "colsum: bad args");
    if (!accumulate) SCAT_CHECK_CUDA(cudaMemsetAsync(out, 0, (size_t)N * sizeof(float), stream));
    const int gx = ceil_div(N, 32);
    int gy = min(ceil_div(M, 8 * 4), max(1, 592 / gx));       // >= 4 rows per thread, ~4 blocks per SM
    if (gy < 1) gy = 1;
    if (x_bf16)
        SCAT_CHECK_CUDA(launch_k(colsum_bf16_kernel, dim3(dim3(gx, gy)), dim3(dim3(32, 8)), 0, stream,
                                 reinterpret_cast<const __nv_bfloat16*>(X), ld, M, N, out));
    else
        SCAT_CHECK_CUDA(launch_k(colsum_kernel, dim3(dim3(gx, gy)), dim3(dim3(32, 8)), 0, stream, X, ld, M, N, out));
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace scat
