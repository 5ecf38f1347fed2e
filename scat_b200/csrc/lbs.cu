// MANO linear blend skinning, models/mano.py:280-391 (rot_pose_beta_to_mesh), fused into one kernel.
//
// The reference issues ~120 small ATen ops per call and materialises posedirs.repeat(B,...) (1.26 MB per
// sample, mano.py:296-300).  Here a CTA owns a group of 8 samples: a set-up phase computes the 16
// Rodrigues rotations, pose-blend weights, regressed joints, the kinematic chain G_i and the skinning
// matrices A_i in shared memory; the vertex phase then sweeps the 778 vertices, one per thread, reading the
// (pre-transposed, vertex-contiguous, L2-resident) blend-shape tables coalesced and re-using every table
// element for all 8 samples from registers.  Work is ~1.2 MFLOP per 9.8 KB of output, so the kernel is
// fp32-ALU bound, not HBM bound.
//
// derived buffer (scat_lbs_prepare), vertex dimension padded to VP = 784:
//   J_template[16*3] | J_shapedirs[16*3*10] | vt_t[3][VP] | sd_t[10][3][VP] | pd_t[135][3][VP] | w_t[16][VP]
#include <stdlib.h>

#include "kernels.h"

namespace scat {
namespace {

constexpr int NV = 778, NJ = 16, NB = 10, NPW = 135, VP = 784;
constexpr int LBS_S = 8;           // samples per CTA
constexpr int LBS_THREADS = 256;
constexpr int OFF_JT = 0, OFF_JS = OFF_JT + NJ * 3, OFF_VT = 528 /* 16*3 + 16*3*10 */, OFF_SD = OFF_VT + 3 * VP,
              OFF_PD = OFF_SD + NB * 3 * VP, OFF_W = OFF_PD + NPW * 3 * VP, DERIVED_FLOATS = OFF_W + NJ * VP;
__constant__ int c_parent[NJ] = {-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14};   // mano.py:221-223
__constant__ int c_tips[5] = {320, 443, 671, 554, 744};                                    // mano.py:373-377

// R = I + sin(t) S(n) + (1 - cos(t)) S(n)^2, n = r/t; Taylor form only where t < 1e-30 (mano.py:236-267)
__device__ void rodrigues(float rx, float ry, float rz, float* R) {
    const float t2 = rx * rx + ry * ry + rz * rz;
    const float t = sqrtf(t2);
    float a, b, nx, ny, nz;
    if (t < 1e-30f) {
        a = 1.0f - t2 / 6.0f; b = 0.5f - t2 / 24.0f; nx = rx; ny = ry; nz = rz;
    } else {
        a = sinf(t); b = 1.0f - cosf(t); nx = rx / t; ny = ry / t; nz = rz / t;
    }
    // S = [[0,-nz,ny],[nz,0,-nx],[-ny,nx,0]];  S^2 = n n^T - |n|^2 I
    const float nn = nx * nx + ny * ny + nz * nz;
    R[0] = 1.0f + b * (nx * nx - nn); R[1] = -a * nz + b * nx * ny;     R[2] = a * ny + b * nx * nz;
    R[3] = a * nz + b * nx * ny;      R[4] = 1.0f + b * (ny * ny - nn); R[5] = -a * nx + b * ny * nz;
    R[6] = -a * ny + b * nx * nz;     R[7] = a * nx + b * ny * nz;      R[8] = 1.0f + b * (nz * nz - nn);
}

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

struct SampleSetup {
    float pw[NPW];          // pose blend weights (R_i - I), i = 1..15, row-major (mano.py:270-277)
    float beta[NB];
    float A[NJ][12];        // skinning matrices, rows of [R | t] (mano.py:331-337)
    float Rg[9];            // global rotation (mano.py:351)
    float root[3];          // rotated joint 1 (mano.py:386)
    float Jtr[NJ][3];       // chain translations before the global rotation
};

__global__ void __launch_bounds__(LBS_THREADS)
lbs_fwd_kernel(const float* __restrict__ derived, const float* __restrict__ hands_mean, const float* __restrict__ rots,
               const float* __restrict__ poses, const float* __restrict__ betas, float* __restrict__ out, int B) {
    pdl_sync();
    __shared__ SampleSetup S[LBS_S];
    __shared__ float Rl[LBS_S][NJ][9];
    __shared__ float Jp[LBS_S][NJ][3];
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * LBS_S;
    const int ns = min(LBS_S, B - b0);

    // ---- set-up phase -----------------------------------------------------------------------
    for (int e = tid; e < LBS_S * NJ; e += LBS_THREADS) {       // local rotations + pose weights
        const int s = e / NJ, i = e % NJ;
        float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (s < ns) {
            if (i == 0) {
                // local root rotation is forced to 0 (mano.py:234,286) -> Taylor branch -> identity
                rodrigues(0.f, 0.f, 0.f, R);
            } else {
                const float* ps = poses + (long long)(b0 + s) * 45 + (i - 1) * 3;
                const float* hm = hands_mean + (i - 1) * 3;
                rodrigues(hm[0] + ps[0], hm[1] + ps[1], hm[2] + ps[2], R);   // no PCA (mano.py:284)
#pragma unroll
                for (int q = 0; q < 9; ++q) S[s].pw[(i - 1) * 9 + q] = R[q] - ((q == 0 || q == 4 || q == 8) ? 1.0f : 0.0f);
            }
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) Rl[s][i][q] = R[q];
    }
    for (int e = tid; e < LBS_S * NB; e += LBS_THREADS) {
        const int s = e / NB, k = e % NB;
        S[s].beta[k] = s < ns ? betas[(long long)(b0 + s) * NB + k] : 0.f;
    }
    if (tid < LBS_S) {
        float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (tid < ns) rodrigues(rots[(long long)(b0 + tid) * 3], rots[(long long)(b0 + tid) * 3 + 1], rots[(long long)(b0 + tid) * 3 + 2], R);
#pragma unroll
        for (int q = 0; q < 9; ++q) S[tid].Rg[q] = R[q];
    }
    __syncthreads();
    for (int e = tid; e < LBS_S * NJ * 3; e += LBS_THREADS) {   // J = Jreg v_shaped (mano.py:302-304)
        const int s = e / (NJ * 3), jc = e % (NJ * 3);
        float v = derived[OFF_JT + jc];
#pragma unroll
        for (int k = 0; k < NB; ++k) v = fmaf(derived[OFF_JS + jc * NB + k], S[s].beta[k], v);
        Jp[s][jc / 3][jc % 3] = v;
    }
    __syncthreads();
    if (tid < LBS_S) {                                          // kinematic chain (mano.py:318-337)
        const int s = tid;
        float G[NJ][12];
#pragma unroll 1
        for (int i = 0; i < NJ; ++i) {
            float L[12];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) L[r * 4 + c] = Rl[s][i][r * 3 + c];
                L[r * 4 + 3] = (i == 0) ? Jp[s][0][r] : Jp[s][i][r] - Jp[s][c_parent[i]][r];
            }
            if (i == 0) {
#pragma unroll
                for (int q = 0; q < 12; ++q) G[0][q] = L[q];
            } else {
                const float* Pm = G[c_parent[i]];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float v = Pm[r * 4 + 0] * L[0 * 4 + c] + Pm[r * 4 + 1] * L[1 * 4 + c] + Pm[r * 4 + 2] * L[2 * 4 + c];
                        if (c == 3) v += Pm[r * 4 + 3];
                        G[i][r * 4 + c] = v;
                    }
                }
            }
        }
#pragma unroll 1
        for (int i = 0; i < NJ; ++i) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float gj = G[i][r * 4 + 0] * Jp[s][i][0] + G[i][r * 4 + 1] * Jp[s][i][1] + G[i][r * 4 + 2] * Jp[s][i][2];
                S[s].A[i][r * 4 + 0] = G[i][r * 4 + 0];
                S[s].A[i][r * 4 + 1] = G[i][r * 4 + 1];
                S[s].A[i][r * 4 + 2] = G[i][r * 4 + 2];
                S[s].A[i][r * 4 + 3] = G[i][r * 4 + 3] - gj;
                S[s].Jtr[i][r] = G[i][r * 4 + 3];
            }
        }
        const float* Rg = S[s].Rg;
#pragma unroll
        for (int r = 0; r < 3; ++r)
            S[s].root[r] = Rg[r * 3 + 0] * S[s].Jtr[1][0] + Rg[r * 3 + 1] * S[s].Jtr[1][1] + Rg[r * 3 + 2] * S[s].Jtr[1][2];
    }
    __syncthreads();
    // chain joints 0..15: rotate, subtract root (mano.py:383-388)
    for (int e = tid; e < ns * NJ * 3; e += LBS_THREADS) {
        const int s = e / (NJ * 3), j = (e / 3) % NJ, r = e % 3;
        const float* Rg = S[s].Rg;
        float v = Rg[r * 3 + 0] * S[s].Jtr[j][0] + Rg[r * 3 + 1] * S[s].Jtr[j][1] + Rg[r * 3 + 2] * S[s].Jtr[j][2] - S[s].root[r];
        if (j == 1) v = 0.f;   // Jtr[:,1] - root is exactly zero in the reference
        out[((long long)(b0 + s) * 799 + j) * 3 + r] = v;
    }

    // ---- vertex phase -----------------------------------------------------------------------
    const float* vt_t = derived + OFF_VT;
    const float* sd_t = derived + OFF_SD;
    const float* pd_t = derived + OFF_PD;
    const float* w_t = derived + OFF_W;
    for (int v = tid; v < NV; v += LBS_THREADS) {
        float vp[LBS_S][3];
        {
            const float m0 = vt_t[v], m1 = vt_t[VP + v], m2 = vt_t[2 * VP + v];
#pragma unroll
            for (int s = 0; s < LBS_S; ++s) { vp[s][0] = m0; vp[s][1] = m1; vp[s][2] = m2; }
        }
#pragma unroll 2
        for (int k = 0; k < NB; ++k) {                          // shape blend shapes (mano.py:288-292)
            const float d0 = sd_t[(k * 3 + 0) * VP + v], d1 = sd_t[(k * 3 + 1) * VP + v], d2 = sd_t[(k * 3 + 2) * VP + v];
#pragma unroll
            for (int s = 0; s < LBS_S; ++s) {
                const float bk = S[s].beta[k];
                vp[s][0] = fmaf(d0, bk, vp[s][0]); vp[s][1] = fmaf(d1, bk, vp[s][1]); vp[s][2] = fmaf(d2, bk, vp[s][2]);
            }
        }
#pragma unroll 3
        for (int k = 0; k < NPW; ++k) {                         // pose blend shapes (mano.py:296-300)
            const float d0 = pd_t[(k * 3 + 0) * VP + v], d1 = pd_t[(k * 3 + 1) * VP + v], d2 = pd_t[(k * 3 + 2) * VP + v];
#pragma unroll
            for (int s = 0; s < LBS_S; ++s) {
                const float wk = S[s].pw[k];
                vp[s][0] = fmaf(d0, wk, vp[s][0]); vp[s][1] = fmaf(d1, wk, vp[s][1]); vp[s][2] = fmaf(d2, wk, vp[s][2]);
            }
        }
        float wj[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) wj[j] = w_t[j * VP + v];
        int tip = -1;
#pragma unroll
        for (int q = 0; q < 5; ++q) if (c_tips[q] == v) tip = q;
#pragma unroll
        for (int s = 0; s < LBS_S; ++s) {                       // skinning (mano.py:339-348)
            if (s >= ns) break;
            float T[12];
#pragma unroll
            for (int q = 0; q < 12; ++q) T[q] = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
#pragma unroll
                for (int q = 0; q < 12; ++q) T[q] = fmaf(wj[j], S[s].A[j][q], T[q]);
            }
            float x[3];
#pragma unroll
            for (int r = 0; r < 3; ++r)
                x[r] = T[r * 4 + 0] * vp[s][0] + T[r * 4 + 1] * vp[s][1] + T[r * 4 + 2] * vp[s][2] + T[r * 4 + 3];
            const float* Rg = S[s].Rg;
            float y[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) y[r] = Rg[r * 3 + 0] * x[0] + Rg[r * 3 + 1] * x[1] + Rg[r * 3 + 2] * x[2] - S[s].root[r];
            float* o = out + ((long long)(b0 + s) * 799 + 21 + v) * 3;
            o[0] = y[0]; o[1] = y[1]; o[2] = y[2];
            if (tip >= 0) {                                     // fingertips from the mesh (mano.py:373-377)
                float* oj = out + ((long long)(b0 + s) * 799 + 16 + tip) * 3;
                oj[0] = y[0]; oj[1] = y[1]; oj[2] = y[2];
            }
        }
    }
}


// ---------------------------------------------------------------------------------------------------------------
// EXPERIMENTAL, off by default (SCAT_LBS_V2="S,NVT" selects it; unvalidated on hardware at the end of round 1, see
// DESIGN.md section 9).  Same arithmetic as lbs_fwd_kernel, different blocking: S samples per CTA (8, 16 or 32) so each
// 1.27 MB pass over the blend-shape tables in L2 serves more samples (159 KB of L2 traffic per sample at S = 8), NVT
// vertices per thread, and the per-sample operands (pose weights, betas, skinning matrices) stored sample-minor /
// 16-byte aligned in shared memory so they are fetched with LDS.128 instead of one LDS per FMA triple.
template <int S>
struct LbsSmemV2 {
    float A[S][NJ][12];          // 48-byte rows: three float4
    float pwT[NPW][S];           // sample-minor
    float betaT[NB][S];
    float Rg[S][9];
    float root[S][3];
    float Jtr[S][NJ][3];
    float Rl[S][NJ][9];
    float Jp[S][NJ][3];
};

template <int S, int NVT>
__device__ __forceinline__ void lbs_vertices_v2(const LbsSmemV2<S>& sm, const float* __restrict__ derived, const int (&vid)[NVT],
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
lbs_fwd_v2_kernel(const float* __restrict__ derived, const float* __restrict__ hands_mean, const float* __restrict__ rots,
                  const float* __restrict__ poses, const float* __restrict__ betas, float* __restrict__ out, int B) {
    pdl_sync();
    extern __shared__ __align__(16) unsigned char lbs_smem_raw[];
    LbsSmemV2<S>& sm = *reinterpret_cast<LbsSmemV2<S>*>(lbs_smem_raw);
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * S;
    const int ns = min(S, B - b0);

    // ---- set-up phase: as lbs_fwd_kernel, stores re-laid out ----
    for (int e = tid; e < S * NJ; e += LBS_THREADS) {
        const int s = e / NJ, i = e % NJ;
        float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (s < ns) {
            if (i == 0) {
                rodrigues(0.f, 0.f, 0.f, R);
            } else {
                const float* ps = poses + (long long)(b0 + s) * 45 + (i - 1) * 3;
                const float* hm = hands_mean + (i - 1) * 3;
                rodrigues(hm[0] + ps[0], hm[1] + ps[1], hm[2] + ps[2], R);
            }
        }
        if (i > 0) {
#pragma unroll
            for (int q = 0; q < 9; ++q)
                sm.pwT[(i - 1) * 9 + q][s] = s < ns ? R[q] - ((q == 0 || q == 4 || q == 8) ? 1.0f : 0.0f) : 0.f;
        }
#pragma unroll
        for (int q = 0; q < 9; ++q) sm.Rl[s][i][q] = R[q];
    }
    for (int e = tid; e < S * NB; e += LBS_THREADS) {
        const int s = e / NB, k = e % NB;
        sm.betaT[k][s] = s < ns ? betas[(long long)(b0 + s) * NB + k] : 0.f;
    }
    if (tid < S) {
        float R[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
        if (tid < ns) rodrigues(rots[(long long)(b0 + tid) * 3], rots[(long long)(b0 + tid) * 3 + 1], rots[(long long)(b0 + tid) * 3 + 2], R);
#pragma unroll
        for (int q = 0; q < 9; ++q) sm.Rg[tid][q] = R[q];
    }
    __syncthreads();
    for (int e = tid; e < S * NJ * 3; e += LBS_THREADS) {
        const int s = e / (NJ * 3), jc = e % (NJ * 3);
        float v = derived[OFF_JT + jc];
#pragma unroll
        for (int k = 0; k < NB; ++k) v = fmaf(derived[OFF_JS + jc * NB + k], sm.betaT[k][s], v);
        sm.Jp[s][jc / 3][jc % 3] = v;
    }
    __syncthreads();
    if (tid < S) {
        const int s = tid;
        float G[NJ][12];
#pragma unroll 1
        for (int i = 0; i < NJ; ++i) {
            float L[12];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) L[r * 4 + c] = sm.Rl[s][i][r * 3 + c];
                L[r * 4 + 3] = (i == 0) ? sm.Jp[s][0][r] : sm.Jp[s][i][r] - sm.Jp[s][c_parent[i]][r];
            }
            if (i == 0) {
#pragma unroll
                for (int q = 0; q < 12; ++q) G[0][q] = L[q];
            } else {
                const float* Pm = G[c_parent[i]];
#pragma unroll
                for (int r = 0; r < 3; ++r) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float v = Pm[r * 4 + 0] * L[0 * 4 + c] + Pm[r * 4 + 1] * L[1 * 4 + c] + Pm[r * 4 + 2] * L[2 * 4 + c];
                        if (c == 3) v += Pm[r * 4 + 3];
                        G[i][r * 4 + c] = v;
                    }
                }
            }
        }
#pragma unroll 1
        for (int i = 0; i < NJ; ++i) {
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                const float gj = G[i][r * 4 + 0] * sm.Jp[s][i][0] + G[i][r * 4 + 1] * sm.Jp[s][i][1] + G[i][r * 4 + 2] * sm.Jp[s][i][2];
                sm.A[s][i][r * 4 + 0] = G[i][r * 4 + 0];
                sm.A[s][i][r * 4 + 1] = G[i][r * 4 + 1];
                sm.A[s][i][r * 4 + 2] = G[i][r * 4 + 2];
                sm.A[s][i][r * 4 + 3] = G[i][r * 4 + 3] - gj;
                sm.Jtr[s][i][r] = G[i][r * 4 + 3];
            }
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
            sm.root[s][r] = sm.Rg[s][r * 3 + 0] * sm.Jtr[s][1][0] + sm.Rg[s][r * 3 + 1] * sm.Jtr[s][1][1] + sm.Rg[s][r * 3 + 2] * sm.Jtr[s][1][2];
    }
    __syncthreads();
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
        if (any) lbs_vertices_v2<S, NVT>(sm, derived, vid, ns, b0, out);
    }
}

template <int S, int NVT>
int launch_lbs_v2(const float* derived, const float* hands_mean, const float* rots, const float* poses, const float* betas,
                  float* out, int B, cudaStream_t stream) {
    const size_t smem = sizeof(LbsSmemV2<S>);
    SCAT_ENSURE_SMEM((lbs_fwd_v2_kernel<S, NVT>), smem);
    SCAT_CHECK_CUDA(launch_k(lbs_fwd_v2_kernel<S, NVT>, dim3(ceil_div(B, S)), dim3(LBS_THREADS), smem, stream, derived, hands_mean,
                             rots, poses, betas, out, B));
    SCAT_CHECK_LAUNCH();
    return 0;
}

// SCAT_LBS_V2 = "S,NVT" with S in {8, 16, 32} and NVT in {1, 2}; anything else (or unset) keeps lbs_fwd_kernel
int lbs_v2_choice() {
    static int choice = -1;
    if (choice < 0) {
        choice = 0;
        const char* e = getenv("SCAT_LBS_V2");
        int s = 0, n = 0;
        if (e && sscanf(e, "%d,%d", &s, &n) == 2 && (s == 8 || s == 16 || s == 32) && (n == 1 || n == 2)) choice = s * 10 + n;
    }
    return choice;
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
    switch (lbs_v2_choice()) {          // experimental blockings, off unless SCAT_LBS_V2 is set
        case 81: return launch_lbs_v2<8, 1>(derived, hands_mean, rots, poses, betas, out, B, stream);
        case 82: return launch_lbs_v2<8, 2>(derived, hands_mean, rots, poses, betas, out, B, stream);
        case 161: return launch_lbs_v2<16, 1>(derived, hands_mean, rots, poses, betas, out, B, stream);
        case 162: return launch_lbs_v2<16, 2>(derived, hands_mean, rots, poses, betas, out, B, stream);
        case 321: return launch_lbs_v2<32, 1>(derived, hands_mean, rots, poses, betas, out, B, stream);
        default: break;
    }
    SCAT_CHECK_CUDA(launch_k(lbs_fwd_kernel, dim3(ceil_div(B, LBS_S)), dim3(LBS_THREADS), 0, stream, derived, hands_mean, rots, poses, betas, out, B));
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
