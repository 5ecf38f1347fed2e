// MANO linear blend skinning with the blend shapes on the tensor cores (models/mano.py:280-391).
//
// 0.65 of the 1.19 MFLOP per sample are the two blend-shape contractions v_shaped = mu + S beta (mano.py:288-292) and
// v_posed = v_shaped + P (R - I) (mano.py:296-300): one dense product  corr[b, 3v+c] = sum_k U[b,k] D[3v+c,k]  with
// U = [beta (10) | pose weights (135)] and D = [shapedirs | posedirs] (2334 x 145).  Here it runs on the tcgen05 GEMM
// kernel (gemm_tc.cu) at fp32 grade: both operands are split into TF32 hi + lo parts stacked along K,
//   U' = [U_hi | U_lo | U_hi],  D' = [D_hi | D_hi | D_lo]   ->   U_hi D_hi + U_lo D_hi + U_hi D_lo   (K' = 3 x 148),
// every value exactly representable on the tensor core, fp32 accumulation in TMEM (the dropped lo x lo term is 2^-22).
// A call is cut into chunks whose intermediates (U' and the [chunk, 2336] corrections, ~10 KB per sample) stay in the
// 126 MB L2:  set-up kernel (Rodrigues, chain, skinning matrices, U')  ->  GEMM  ->  skinning kernel (thread per vertex:
// T = sum_j w_vj A_j, x = T (v_posed, 1), global rotation, root).  The FFMA kernel of lbs.cu stays as the fp32 form.
#include "lbs_common.cuh"
#include <mutex>

namespace scat {

size_t lbs_derived_floats();

namespace {

constexpr int TC_KP = 148;                 // 145 blend shapes padded to 16-byte rows
constexpr int TC_K = 3 * TC_KP;            // stacked hi / lo / hi
constexpr int TC_LD = 448;                 // row pitch of U' and D' in floats: 128-byte rows (at pitch 444 = 1776 bytes the GEMM
                                           // takes 68.8 us per 8192 samples, at 448 44.7 us: tools/gemm_pitch_probe.py)
constexpr int TC_N = NV * 3;               // 2334 vertex coordinates, n = 3 v + c
constexpr int TC_LDC = 2336;
constexpr int TC_S = 16;                   // samples per CTA in the set-up kernel
constexpr int TC_SETUP_FLOATS = 672;                          // per sample: U' (448), A (192), (Rg | root) (12), padded to 128 bytes
static_assert(TC_SETUP_FLOATS >= TC_LD + NJ * 12 + 12 && TC_SETUP_FLOATS % 32 == 0, "set-up set layout");
// scratch floats per sample: TWO set-up sets (the set-up of chunk c + 1 runs on a second stream beside the GEMM and the
// skinning of chunk c) + the corrections
constexpr int TC_PER_SAMPLE = 2 * TC_SETUP_FLOATS + TC_LDC;

// second stream + events of scat_lbs_fwd_tc, per device (created on first use; a call holds the mutex while it enqueues)
struct LbsSide {
    cudaStream_t s = nullptr;
    cudaEvent_t ev[8];
    int next = 0;
    bool ready = false;
    std::mutex mu;
};
LbsSide g_lbs_side[16];
std::mutex g_lbs_side_create;
LbsSide* lbs_side() {
    static const int on = [] { const char* e = getenv("SCAT_LBS_PIPELINE"); return (e && e[0] == '0') ? 0 : 1; }();
    int dev = 0;
    if (!on || cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    LbsSide& sd = g_lbs_side[dev];
    std::lock_guard<std::mutex> lock(g_lbs_side_create);
    if (!sd.ready) {
        if (cudaStreamCreateWithFlags(&sd.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (int i = 0; i < 8; ++i)
            if (cudaEventCreateWithFlags(&sd.ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        sd.ready = true;
    }
    return &sd;
}

// D' [2334, 444] at row pitch 448: rows n = 3 v + c, columns [D_hi | D_hi | D_lo], D = [shapedirs[v,c,:] | posedirs[v,c,:] | 0 0 0]
__global__ void lbs_tc_prepare_kernel(const float* __restrict__ shapedirs, const float* __restrict__ posedirs,
                                      float* __restrict__ Dst) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TC_N * TC_K) return;
    const int n = i / TC_K, kk = i - n * TC_K, third = kk / TC_KP, k = kk - third * TC_KP;
    const float v = k < NB ? shapedirs[n * NB + k] : (k < NB + NPW ? posedirs[n * NPW + (k - NB)] : 0.f);
    const float hi = round_tf32(v);
    Dst[(long long)n * TC_LD + kk] = third == 2 ? round_tf32(v - hi) : hi;       // (columns 444..447 of a row are never read)
}

// per sample: U' row, skinning matrices A_j (rows of [R | t]), global rotation and root; joints 0..15 go straight out
__global__ void __launch_bounds__(LBS_THREADS)
lbs_tc_setup_kernel(const float* __restrict__ derived, const float* __restrict__ hands_mean, const float* __restrict__ rots,
                    const float* __restrict__ poses, const float* __restrict__ betas, float* __restrict__ U,
                    float* __restrict__ Aout, float* __restrict__ Rr, float* __restrict__ out, int b_first, int n_chunk) {
    pdl_sync();
    extern __shared__ __align__(16) unsigned char lbs_tc_raw[];
    LbsSetup<TC_S>& sm = *reinterpret_cast<LbsSetup<TC_S>*>(lbs_tc_raw);
    const int tid = threadIdx.x;
    const int l0 = blockIdx.x * TC_S;                       // first sample of this CTA inside the chunk
    const int ns = min(TC_S, n_chunk - l0);
    const int b0 = b_first + l0;
    lbs_setup<TC_S>(sm, derived, hands_mean, rots, poses, betas, b0, ns);
    for (int e = tid; e < ns * NJ * 3; e += LBS_THREADS) {  // chain joints: rotate, subtract root (mano.py:383-388)
        const int s = e / (NJ * 3), j = (e / 3) % NJ, r = e % 3;
        const float* Rg = sm.Rg[s];
        float v = Rg[r * 3 + 0] * sm.Jtr[s][j][0] + Rg[r * 3 + 1] * sm.Jtr[s][j][1] + Rg[r * 3 + 2] * sm.Jtr[s][j][2] - sm.root[s][r];
        if (j == 1) v = 0.f;
        out[((long long)(b0 + s) * 799 + j) * 3 + r] = v;
    }
    for (int e = tid; e < ns * TC_KP; e += LBS_THREADS) {   // U' = [hi | lo | hi]
        const int s = e / TC_KP, k = e % TC_KP;
        const float v = k < NB ? sm.betaT[k][s] : (k < NB + NPW ? sm.pwT[k - NB][s] : 0.f);
        const float hi = round_tf32(v), lo = round_tf32(v - hi);
        float* u = U + (long long)(l0 + s) * TC_LD;
        u[k] = hi; u[TC_KP + k] = lo; u[2 * TC_KP + k] = hi;
    }
    for (int e = tid; e < ns * NJ * 12; e += LBS_THREADS) {
        const int s = e / (NJ * 12), q = e % (NJ * 12);
        Aout[(long long)(l0 + s) * NJ * 12 + q] = (&sm.A[s][0][0])[q];
    }
    for (int e = tid; e < ns * 12; e += LBS_THREADS) {
        const int s = e / 12, q = e % 12;
        Rr[(long long)(l0 + s) * 12 + q] = q < 9 ? sm.Rg[s][q] : sm.root[s][q - 9];
    }
}

// one CTA per sample, FOUR consecutive vertices per thread: skinning (mano.py:339-348), global rotation and root (:382-388),
// fingertips (:373-377).  A thread per vertex was bound by the load/store unit (per joint one weight load and three
// broadcast LDS.128 of A_j for 12 FMAs: ncu 134 us per 8192 samples, 25 % of the FMA peak); with four vertices a joint
// costs one LDG.128 of weights and the same three LDS.128 for 48 FMAs.  Per vertex the arithmetic (and its order) is
// unchanged.  Results leave through a shared-memory row so that the 9.3 KB of a sample's vertices are written coalesced.
constexpr int SKIN_VPT = 4;
static_assert(VP % SKIN_VPT == 0 && (OFF_VT % 4) == 0 && (OFF_W % 4) == 0 && (TC_LDC % 4) == 0, "float4 table reads");
constexpr int SKIN_THREADS = 224;             // 7 warps: 195 threads own vertices
static_assert((NV + SKIN_VPT - 1) / SKIN_VPT <= SKIN_THREADS, "one pass over the vertices");
__global__ void __launch_bounds__(SKIN_THREADS, 3)
lbs_tc_skin_kernel(const float* __restrict__ derived, const float* __restrict__ corr, const float* __restrict__ Ain,
                   const float* __restrict__ Rr, float* __restrict__ out, int b_first) {
    pdl_sync();
    __shared__ __align__(16) float A[NJ][12];
    __shared__ float R[12];
    __shared__ __align__(16) float stage[NV * 3 + 2];
    const int l = blockIdx.x, tid = threadIdx.x;
    for (int e = tid; e < NJ * 12; e += SKIN_THREADS) (&A[0][0])[e] = Ain[(long long)l * NJ * 12 + e];
    if (tid < 12) R[tid] = Rr[(long long)l * 12 + tid];
    __syncthreads();
    const float* vt_t = derived + OFF_VT;
    const float* w_t = derived + OFF_W;
    const int v0 = tid * SKIN_VPT;
    if (v0 < NV) {
        const float* cr = corr + (long long)l * TC_LDC + 3 * v0;       // 12 corrections of vertices v0 .. v0 + 3
        // (read after the joint loop, so that they are not live across it: three CTAs per SM; requested now)
        asm volatile("prefetch.global.L1 [%0];" ::"l"(cr));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(cr + 8));
        float T[SKIN_VPT][12];
#pragma unroll
        for (int u = 0; u < SKIN_VPT; ++u)
#pragma unroll
            for (int q = 0; q < 12; ++q) T[u][q] = 0.f;
#pragma unroll 2
        for (int j = 0; j < NJ; ++j) {
            const float4 w4 = __ldg(reinterpret_cast<const float4*>(w_t + j * VP + v0));
            const float w[4] = {w4.x, w4.y, w4.z, w4.w};
            const float4* arow = reinterpret_cast<const float4*>(A[j]);
            const float4 a0 = arow[0], a1 = arow[1], a2 = arow[2];
            const float a[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
            // packed fp32 FMA (FFMA2, sm_100): two of a vertex's 12 accumulators per instruction, each lane an ordinary
            // fma.rn -- the same values as fmaf, half the issue slots of a kernel that ncu shows issue-bound (66 % busy)
#pragma unroll
            for (int u = 0; u < SKIN_VPT; ++u)
#pragma unroll
                for (int q = 0; q < 12; q += 2) ffma2(T[u][q], T[u][q + 1], w[u], a[q], a[q + 1]);
        }
        float c[12];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (3 * v0 + 4 * k < TC_LDC) t4 = __ldg(reinterpret_cast<const float4*>(cr) + k);     // (row end: vertices >= NV)
            c[4 * k] = t4.x; c[4 * k + 1] = t4.y; c[4 * k + 2] = t4.z; c[4 * k + 3] = t4.w;
        }
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(vt_t + v0));
        const float4 t1 = __ldg(reinterpret_cast<const float4*>(vt_t + VP + v0));
        const float4 t2 = __ldg(reinterpret_cast<const float4*>(vt_t + 2 * VP + v0));
        const float tx[4] = {t0.x, t0.y, t0.z, t0.w}, ty[4] = {t1.x, t1.y, t1.z, t1.w}, tz[4] = {t2.x, t2.y, t2.z, t2.w};
#pragma unroll
        for (int u = 0; u < SKIN_VPT; ++u) {
            if (v0 + u < NV) {
                const float p0 = tx[u] + c[3 * u], p1 = ty[u] + c[3 * u + 1], p2 = tz[u] + c[3 * u + 2];
                float x[3];
#pragma unroll
                for (int r = 0; r < 3; ++r) x[r] = T[u][r * 4 + 0] * p0 + T[u][r * 4 + 1] * p1 + T[u][r * 4 + 2] * p2 + T[u][r * 4 + 3];
#pragma unroll
                for (int r = 0; r < 3; ++r)
                    stage[3 * (v0 + u) + r] = R[r * 3 + 0] * x[0] + R[r * 3 + 1] * x[1] + R[r * 3 + 2] * x[2] - R[9 + r];
            }
        }
    }
    __syncthreads();
    float* o_base = out + (long long)(b_first + l) * 799 * 3;
    for (int i = tid; i < NV * 3; i += SKIN_THREADS) o_base[21 * 3 + i] = stage[i];
    if (tid < 15) {
        const int q = tid / 3, r = tid - 3 * q;
        o_base[(16 + q) * 3 + r] = stage[c_tips[q] * 3 + r];
    }
}

}  // namespace

size_t lbs_tc_table_floats() { return (size_t)TC_N * TC_LD; }

}  // namespace scat

using namespace scat;

extern "C" {

size_t scat_lbs_tc_table_floats(void) { return lbs_tc_table_floats(); }
size_t scat_lbs_tc_scratch_floats(int32_t batch) {
    // ~14 KB per sample, 74 MB of corrections through L2 per chunk.  7936 samples = 62 row tiles x 19 column tiles = 1178
    // GEMM tiles = 3.98 rounds of the 296 persistent CTAs (8192 samples are 1216 tiles = 4.1 rounds, i.e. five)
    const long long chunk = batch < 7936 ? (batch > 0 ? batch : 1) : 7936;
    return (size_t)chunk * TC_PER_SAMPLE;
}

int scat_lbs_tc_prepare(const float* shapedirs, const float* posedirs, float* table, void* stream) {
    SCAT_REQUIRE(shapedirs && posedirs && table, kErrBadArg, "lbs_tc_prepare: null");
    const int n = TC_N * TC_K;
    SCAT_CHECK_CUDA(launch_k(lbs_tc_prepare_kernel, dim3(ceil_div(n, 256)), dim3(256), 0, (cudaStream_t)stream, shapedirs, posedirs, table));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int scat_lbs_fwd_tc(const float* derived, const float* table, const float* hands_mean, const float* rots, const float* poses,
                    const float* betas, float* out, int32_t batch, float* scratch, size_t scratch_floats, void* stream) {
    SCAT_REQUIRE(derived && table && hands_mean && rots && poses && betas && out && scratch && batch > 0, kErrBadArg,
                 "lbs_fwd_tc: bad args");
    SCAT_REQUIRE(((uintptr_t)scratch & 15) == 0 && ((uintptr_t)table & 15) == 0, kErrBadArg, "lbs_fwd_tc: scratch / table must be 16-byte aligned");
    const long long chunk_max = (long long)(scratch_floats / TC_PER_SAMPLE);
    SCAT_REQUIRE(chunk_max >= 1, kErrWorkspace, "lbs_fwd_tc: scratch holds no sample (%zu floats, %d per sample)", scratch_floats,
                 TC_PER_SAMPLE);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t smem = sizeof(LbsSetup<TC_S>);
    SCAT_ENSURE_SMEM(lbs_tc_setup_kernel, smem);
    // chunk c: set-up (Rodrigues, chain, U') -> GEMM -> skinning.  The set-up of chunk c + 1 only needs its own buffers, so
    // it runs on a second stream beside the GEMM / skinning of chunk c (two set-up sets alternate); the main stream waits
    // for it just before the GEMM that reads it.  Everything is joined again before the call returns (capturable).
    const long long chunk = batch < chunk_max ? batch : chunk_max;
    float* set_base[2] = {scratch, scratch + (size_t)chunk * TC_SETUP_FLOATS};
    float* corr = scratch + 2 * (size_t)chunk * TC_SETUP_FLOATS;         // [chunk, 2336]
    const int n_chunks = (int)((batch + chunk - 1) / chunk);
    LbsSide* sd = n_chunks > 1 ? lbs_side() : nullptr;
    std::unique_lock<std::mutex> lock;
    if (sd) lock = std::unique_lock<std::mutex>(sd->mu);
    auto next_event = [&]() { cudaEvent_t e = sd->ev[sd->next]; sd->next = (sd->next + 1) % 8; return e; };
    auto setup = [&](int c, cudaStream_t s) -> int {
        const long long b0 = (long long)c * chunk;
        const int n = (int)(batch - b0 < chunk ? batch - b0 : chunk);
        float* U = set_base[c & 1];
        float* A = U + (size_t)chunk * TC_LD;
        float* Rr = A + (size_t)chunk * NJ * 12;
        SCAT_CHECK_CUDA(launch_k(lbs_tc_setup_kernel, dim3(ceil_div(n, TC_S)), dim3(LBS_THREADS), smem, s, derived, hands_mean, rots,
                                 poses, betas, U, A, Rr, out, (int)b0, n));
        SCAT_CHECK_LAUNCH();
        return 0;
    };
    SCAT_PROPAGATE(setup(0, st));
    cudaEvent_t ready = nullptr;
    for (int c = 0; c < n_chunks; ++c) {
        const long long b0 = (long long)c * chunk;
        const int n = (int)(batch - b0 < chunk ? batch - b0 : chunk);
        float* U = set_base[c & 1];
        float* A = U + (size_t)chunk * TC_LD;
        float* Rr = A + (size_t)chunk * NJ * 12;
        cudaEvent_t ready_next = nullptr;
        if (c + 1 < n_chunks && sd) {
            // set (c + 1) & 1 was last read by the GEMM / skinning of chunk c - 1: everything on `st` so far
            cudaEvent_t e_free = next_event();
            SCAT_CHECK_CUDA(cudaEventRecord(e_free, st));
            SCAT_CHECK_CUDA(cudaStreamWaitEvent(sd->s, e_free, 0));
            SCAT_PROPAGATE(setup(c + 1, sd->s));
            ready_next = next_event();
            SCAT_CHECK_CUDA(cudaEventRecord(ready_next, sd->s));
        }
        if (ready != nullptr) SCAT_CHECK_CUDA(cudaStreamWaitEvent(st, ready, 0));
        GemmArgs g;
        g.A = U; g.sam = TC_LD; g.sak = 1; g.B = table; g.sbn = TC_LD; g.sbk = 1;
        g.C = corr; g.ldc = TC_LDC; g.M = n; g.N = TC_N; g.K = TC_K; g.prerounded = 1;
        SCAT_PROPAGATE(launch_gemm_tc(g, PREC_TF32, st));
        // the skinning kernel re-reads the 50 KB weight table and the 9 KB template per sample: it wants them in L1, not the
        // maximum shared-memory carve-out every other kernel of the library asks for (11 KB of shared memory per CTA)
        ensure_carveout(reinterpret_cast<const void*>(lbs_tc_skin_kernel), 25);
        SCAT_CHECK_CUDA(launch_k(lbs_tc_skin_kernel, dim3(n), dim3(SKIN_THREADS), 0, st, derived, (const float*)corr, (const float*)A,
                                 (const float*)Rr, out, (int)b0));
        SCAT_CHECK_LAUNCH();
        if (c + 1 < n_chunks && !sd) SCAT_PROPAGATE(setup(c + 1, st));
        ready = ready_next;
    }
    return 0;
}

}  // extern "C"
