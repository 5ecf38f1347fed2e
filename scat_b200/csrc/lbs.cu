// MANO linear blend skinning, models/mano.py:280-391 (rot_pose_beta_to_mesh), fused into one kernel.
//
// The reference issues ~120 small ATen ops per call and materialises posedirs.repeat(B,...) (1.26 MB per
// sample, mano.py:296-300).  Here a CTA owns a group of 16 samples: the set-up phase (lbs_common.cuh) computes the 16
// Rodrigues rotations, pose-blend weights, regressed joints, the kinematic chain G_i and the skinning
// matrices A_i in shared memory; the vertex phase then sweeps the 778 vertices, one per thread, reading the
// (pre-transposed, vertex-contiguous, L2-resident) blend-shape tables coalesced and re-using every table
// element for all 16 samples from registers.  Work is ~1.2 MFLOP per 9.8 KB of output, so the kernel is
// fp32-ALU bound, not HBM bound.  Blocking chosen on hardware (round 2, 65,536 samples):
// 8 samples per CTA 3.42 ms, 16 samples 3.34 ms, 32 samples 4.49 ms, two vertices per thread 4.5-5.3 ms.
#include "lbs_common.cuh"

namespace scat {
namespace {

constexpr int LBS_S = 16;          // samples per CTA

__global__ void lbs_prepare_kernel(const float* __restrict__ v_template, const float* __restrict__ shapedirs,
                                   const float* __restrict__ posedirs, const float* __restrict__ J_reg,
                                   const float* __restrict__ weights, float* __restrict__ derived) {
    pdl_sync();
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = DERIVED_FLOATS;
    if (i >= total) return;
    float out = 0.f;
    if (i < OFF_JS) {                       // J_template[j][c] = sum_v Jreg[j,v] v_template[v,c]
        const int j = i / 3, c = i % 3;
        for (int v = 0; v < NV; ++v) out = fmaf(J_reg[j * NV + v], v_template[v * 3 + c], out);
    } else if (i < OFF_VT) {                // J_shapedirs[j][c][k]
        const int r = i - OFF_JS, j = r / 30, c = (r / 10) % 3, k = r % 10;
        for (int v = 0; v < NV; ++v) out = fmaf(J_reg[j * NV + v], shapedirs[(v * 3 + c) * NB + k], out);
    } else if (i < OFF_SD) {
        const int r = i - OFF_VT, c = r / VP, v = r % VP;
        out = v < NV ? v_template[v * 3 + c] : 0.f;
    } else if (i < OFF_PD) {
        const int r = i - OFF_SD, k = r / (3 * VP), c = (r / VP) % 3, v = r % VP;
        out = v < NV ? shapedirs[(v * 3 + c) * NB + k] : 0.f;
    } else if (i < OFF_W) {
        const int r = i - OFF_PD, k = r / (3 * VP), c = (r / VP) % 3, v = r % VP;
        out = v < NV ? posedirs[(v * 3 + c) * NPW + k] : 0.f;
    } else {
        const int r = i - OFF_W, j = r / VP, v = r % VP;
        out = v < NV ? weights[v * NJ + j] : 0.f;
    }
    derived[i] = out;
}

template <int S, int NVT>
__device__ __forceinline__ void lbs_vertices(const LbsSetup<S>& sm, const float* __restrict__ derived, const int (&vid)[NVT],
                                                int ns, int b0, float* __restrict__ out) {
    const float* vt_t = derived + OFF_VT;
    const float* sd_t = derived + OFF_SD;
    const float* pd_t = derived + OFF_PD;
    const float* w_t = derived + OFF_W;
    float vp[NVT][S][3];
#pragma unroll
    for (int t = 0; t < NVT; ++t) {
        const float m0 = vt_t[vid[t]], m1 = vt_t[VP + vid[t]], m2 = vt_t[2 * VP + vid[t]];
#pragma unroll
        for (int s = 0; s < S; ++s) { vp[t][s][0] = m0; vp[t][s][1] = m1; vp[t][s][2] = m2; }
    }
#pragma unroll 1
    for (int k = 0; k < NB + NPW; ++k) {                    // shape then pose blend shapes (tables are adjacent)
        const float* tab = k < NB ? sd_t + (size_t)k * 3 * VP : pd_t + (size_t)(k - NB) * 3 * VP;
        const float4* wrow = reinterpret_cast<const float4*>(k < NB ? sm.betaT[k] : sm.pwT[k - NB]);
        float d[NVT][3];
#pragma unroll
        for (int t = 0; t < NVT; ++t) { d[t][0] = tab[vid[t]]; d[t][1] = tab[VP + vid[t]]; d[t][2] = tab[2 * VP + vid[t]]; }
#pragma unroll
        for (int s4 = 0; s4 < S / 4; ++s4) {
            const float4 w = wrow[s4];
            const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int t = 0; t < NVT; ++t) {
                    vp[t][s4 * 4 + u][0] = fmaf(d[t][0], ws[u], vp[t][s4 * 4 + u][0]);
                    vp[t][s4 * 4 + u][1] = fmaf(d[t][1], ws[u], vp[t][s4 * 4 + u][1]);
                    vp[t][s4 * 4 + u][2] = fmaf(d[t][2], ws[u], vp[t][s4 * 4 + u][2]);
                }
        }
    }
    float wj[NVT][NJ];
    int tip[NVT];
#pragma unroll
    for (int t = 0; t < NVT; ++t) {
#pragma unroll
        for (int j = 0; j < NJ; ++j) wj[t][j] = w_t[j * VP + vid[t]];
        tip[t] = -1;
#pragma unroll
        for (int q = 0; q < 5; ++q) if (c_tips[q] == vid[t]) tip[t] = q;
    }
#pragma unroll
    for (int s = 0; s < S; ++s) {                           // skinning (mano.py:339-348)
        if (s >= ns) break;
        float T[NVT][12];
#pragma unroll
        for (int t = 0; t < NVT; ++t)
#pragma unroll
            for (int q = 0; q < 12; ++q) T[t][q] = 0.f;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
            const float4* arow = reinterpret_cast<const float4*>(sm.A[s][j]);
            const float4 a0 = arow[0], a1 = arow[1], a2 = arow[2];
            const float a[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
#pragma unroll
            for (int t = 0; t < NVT; ++t)
#pragma unroll
                for (int q = 0; q < 12; ++q) T[t][q] = fmaf(wj[t][j], a[q], T[t][q]);
        }
        const float* Rg = sm.Rg[s];
#pragma unroll
        for (int t = 0; t < NVT; ++t) {
            if (vid[t] >= NV) continue;                     // padded lanes of the remainder pass
            float x[3], y[3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
                x[r] = T[t][r * 4 + 0] * vp[t][s][0] + T[t][r * 4 + 1] * vp[t][s][1] + T[t][r * 4 + 2] * vp[t][s][2] + T[t][r * 4 + 3];
#pragma unroll
            for (int r = 0; r < 3; ++r) y[r] = Rg[r * 3 + 0] * x[0] + Rg[r * 3 + 1] * x[1] + Rg[r * 3 + 2] * x[2] - sm.root[s][r];
            float* o = out + ((long long)(b0 + s) * 799 + 21 + vid[t]) * 3;
            o[0] = y[0]; o[1] = y[1]; o[2] = y[2];
            if (tip[t] >= 0) {
                float* oj = out + ((long long)(b0 + s) * 799 + 16 + tip[t]) * 3;
                oj[0] = y[0]; oj[1] = y[1]; oj[2] = y[2];
            }
        }
    }
}

template <int S, int NVT>
__global__ void __launch_bounds__(LBS_THREADS)
lbs_fwd_kernel(const float* __restrict__ derived, const float* __restrict__ hands_mean, const float* __restrict__ rots,
               const float* __restrict__ poses, const float* __restrict__ betas, float* __restrict__ out, int B) {
    pdl_sync();
    extern __shared__ __align__(16) unsigned char lbs_smem_raw[];
    LbsSetup<S>& sm = *reinterpret_cast<LbsSetup<S>*>(lbs_smem_raw);
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * S;
    const int ns = min(S, B - b0);
    lbs_setup<S>(sm, derived, hands_mean, rots, poses, betas, b0, ns);
    // chain joints 0..15: rotate, subtract root (mano.py:383-388)
    for (int e = tid; e < ns * NJ * 3; e += LBS_THREADS) {
        const int s = e / (NJ * 3), j = (e / 3) % NJ, r = e % 3;
        const float* Rg = sm.Rg[s];
        float v = Rg[r * 3 + 0] * sm.Jtr[s][j][0] + Rg[r * 3 + 1] * sm.Jtr[s][j][1] + Rg[r * 3 + 2] * sm.Jtr[s][j][2] - sm.root[s][r];
        if (j == 1) v = 0.f;
        out[((long long)(b0 + s) * 799 + j) * 3 + r] = v;
    }

    // ---- vertex phase: NVT vertices per thread; the table rows are padded to VP = 784, so lanes past 777 read zeros ----
    for (int v0 = tid; v0 < VP; v0 += LBS_THREADS * NVT) {
        int vid[NVT];
#pragma unroll
        for (int t = 0; t < NVT; ++t) vid[t] = min(v0 + t * LBS_THREADS, VP - 1);
        bool any = false;
#pragma unroll
        for (int t = 0; t < NVT; ++t) {
            if (v0 + t * LBS_THREADS >= NV) vid[t] = VP - 1;         // padding column: results discarded
            any |= vid[t] < NV;
        }
        if (any) lbs_vertices<S, NVT>(sm, derived, vid, ns, b0, out);
    }
}


}  // namespace

size_t lbs_derived_floats() { return DERIVED_FLOATS; }

int launch_lbs_prepare_all(const float* v_template, const float* shapedirs, const float* posedirs, const float* J_reg,
                           const float* weights, float* derived, cudaStream_t stream) {
    SCAT_REQUIRE(v_template && shapedirs && posedirs && J_reg && weights && derived, kErrBadArg, "lbs_prepare: null");
    SCAT_CHECK_CUDA(launch_k(lbs_prepare_kernel, dim3(ceil_div(DERIVED_FLOATS, 256)), dim3(256), 0, stream, v_template, shapedirs, posedirs, J_reg, weights,
                                                                          derived));
    SCAT_CHECK_LAUNCH();
    return 0;
}

int launch_lbs_fwd_derived(const float* derived, const float* hands_mean, const float* rots, const float* poses,
                           const float* betas, float* out, int B, cudaStream_t stream) {
    SCAT_REQUIRE(derived && hands_mean && rots && poses && betas && out && B > 0, kErrBadArg, "lbs_fwd: bad args");
    const size_t smem = sizeof(LbsSetup<LBS_S>);
    SCAT_ENSURE_SMEM((lbs_fwd_kernel<LBS_S, 1>), smem);
    SCAT_CHECK_CUDA(launch_k(lbs_fwd_kernel<LBS_S, 1>, dim3(ceil_div(B, LBS_S)), dim3(LBS_THREADS), smem, stream, derived,
                             hands_mean, rots, poses, betas, out, B));
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace scat

extern "C" {
size_t scat_lbs_derived_floats(void) { return scat::lbs_derived_floats(); }
int scat_lbs_prepare(const float* v_template, const float* shapedirs, const float* posedirs, const float* j_regressor,
                     const float* weights, float* derived, void* stream) {
    return scat::launch_lbs_prepare_all(v_template, shapedirs, posedirs, j_regressor, weights, derived, (cudaStream_t)stream);
}
int scat_lbs_fwd(const float* derived, const float* hands_mean, const float* rots, const float* poses,
                 const float* betas, float* out, int32_t batch, void* stream) {
    return scat::launch_lbs_fwd_derived(derived, hands_mean, rots, poses, betas, out, batch, (cudaStream_t)stream);
}
}
