// Shared pieces of the MANO linear-blend-skinning kernels (lbs.cu forward, lbs_bwd.cu backward): asset layout,
// Rodrigues, and the per-CTA set-up phase (local rotations, pose weights, regressed joints, kinematic chain, skinning
// matrices) of models/mano.py:280-337.
#pragma once
#include "kernels.h"

namespace scat {
namespace {

constexpr int NV = 778, NJ = 16, NB = 10, NPW = 135, VP = 784;
constexpr int LBS_THREADS = 256;
// derived buffer (scat_lbs_prepare), vertex dimension padded to VP = 784:
//   J_template[16*3] | J_shapedirs[16*3*10] | vt_t[3][VP] | sd_t[10][3][VP] | pd_t[135][3][VP] | w_t[16][VP]
constexpr int OFF_JT = 0, OFF_JS = OFF_JT + NJ * 3, OFF_VT = 528 /* 16*3 + 16*3*10 */, OFF_SD = OFF_VT + 3 * VP,
              OFF_PD = OFF_SD + NB * 3 * VP, OFF_W = OFF_PD + NPW * 3 * VP, DERIVED_FLOATS = OFF_W + NJ * VP;
__constant__ int c_parent[NJ] = {-1, 0, 1, 2, 0, 4, 5, 0, 7, 8, 0, 10, 11, 0, 13, 14};   // mano.py:221-223
__constant__ int c_tips[5] = {320, 443, 671, 554, 744};                                    // mano.py:373-377


// R = I + sin(t) S(n) + (1 - cos(t)) S(n)^2, n = r/t; Taylor form only where t < 1e-30 (mano.py:236-267)
__device__ inline void rodrigues(float rx, float ry, float rz, float* R) {
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

// Per-CTA state of S samples.  Per-sample operands of the vertex sweep are stored sample-minor / 16-byte aligned so
// that they are fetched with LDS.128 (one load per four samples).
template <int S>
struct LbsSetup {
    float A[S][NJ][12];          // skinning matrices, rows of [R | t] (mano.py:331-337); 48-byte rows: three float4
    float pwT[NPW][S];           // pose blend weights (R_i - I), i = 1..15, row-major (mano.py:270-277), sample-minor
    float betaT[NB][S];
    float Rg[S][9];              // global rotation (mano.py:351)
    float root[S][3];            // rotated joint 1 (mano.py:386)
    float Jtr[S][NJ][3];         // chain translations before the global rotation
    float Rl[S][NJ][9];          // local rotations
    float Jp[S][NJ][3];          // regressed joints of the shaped template (mano.py:302-304)
    float G[S][NJ][12];          // chain transforms (mano.py:318-327), rows of [R | t]
};

// set-up phase of samples b0 .. b0 + ns - 1 by all LBS_THREADS threads of the CTA; ends with a __syncthreads()
template <int S>
__device__ inline void lbs_setup(LbsSetup<S>& sm, const float* __restrict__ derived, const float* __restrict__ hands_mean,
                                 const float* __restrict__ rots, const float* __restrict__ poses,
                                 const float* __restrict__ betas, int b0, int ns) {
    const int tid = threadIdx.x;
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
        float (*G)[12] = sm.G[s];
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
}

}  // namespace
}  // namespace scat
