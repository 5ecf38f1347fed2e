// Backward of MANO linear blend skinning: the reference's rot_pose_beta_to_mesh (models/mano.py:280-391) is plain
// differentiable PyTorch and runs under autograd; this kernel is that vector-Jacobian product written out by hand.
//
//   grad_out[B, 799, 3]  ->  grad_rots[B,3], grad_poses[B,45], grad_betas[B,10]      (the asset tensors are constants)
//
// One CTA owns 8 samples.  Phases (the forward quantities are recomputed, nothing is saved by scat_lbs_fwd):
//   0  set-up as the forward (lbs_common.cuh): local rotations R_i, regressed joints J_i, chain G_i, skinning A_i
//   1  vertex sweep, thread = vertex, 8 samples in registers: recompute v_posed (shape + pose blend shapes) and the skinned
//      point x; gx = Rg^T gy; keep (gx, v_posed) per (sample, vertex) in shared memory; reduce sum gy and sum gy (x) x
//   2  gA[j] = sum_v w[v,j] [gx (x) v_posed | gx]           (warp per (sample, joint), lanes stride the vertices)
//   3  g_vposed = T_v^T gx with T_v = sum_j w[v,j] A_j, written over v_posed in shared memory
//   4  g_poseweights[k] = sum_v posedirs[v,:,k] . g_vposed[v],  g_betas[k] += sum_v shapedirs[v,:,k] . g_vposed[v]
//      (warp per blend shape, every table element re-used for the 8 samples)
//   5  one thread per sample: chain backward (G_i = G_parent [R_i | J_i - J_parent], A_i = [G_i.R | G_i.t - G_i.R J_i]),
//      joints -> betas through J_shapedirs, Rodrigues backward for the 15 local rotations and the global rotation
// Arithmetic is fp32 FFMA like the forward; cost is about three forward passes (the pose blend runs once forward and once
// transposed).
#include "lbs_common.cuh"

namespace scat {
namespace {

constexpr int BW_S = 8;

struct LbsBwdSmem {
    LbsSetup<BW_S> fw;
    float gx[BW_S][NV][3];        // gradient w.r.t. the skinned, un-rotated vertex
    float vp[BW_S][NV][3];        // v_posed, later overwritten by its gradient
    float gA[BW_S][NJ][12];
    float gblend[BW_S][NPW + NB]; // gradient w.r.t. the pose weights (135) and, through the shape dirs, the betas (10)
    float gRg[BW_S][9];           // sum_p gy_p (x) p over all 799 un-rotated points
    float gsum[BW_S][3];          // sum_p gy_p
    float gJtr[BW_S][NJ][3];      // gradient w.r.t. the chain translations (joint outputs 0..15)
};

// gradient of R = rodrigues(r) (lbs_common.cuh, both branches) w.r.t. r, given gR (row-major 3x3)
__device__ void rodrigues_bwd(float rx, float ry, float rz, const float* gR, float* gr) {
    const float t2 = rx * rx + ry * ry + rz * rz;
    const float t = sqrtf(t2);
    // <gR, S(e_k)>: S(e_0) = [[0,0,0],[0,0,-1],[0,1,0]], S(e_1) = [[0,0,1],[0,0,0],[-1,0,0]], S(e_2) = [[0,-1,0],[1,0,0],[0,0,0]]
    const float s0 = gR[7] - gR[5], s1 = gR[2] - gR[6], s2 = gR[3] - gR[1];
    const float tr = gR[0] + gR[4] + gR[8];
    const bool small = t < 1e-30f;
    const float nx = small ? rx : rx / t, ny = small ? ry : ry / t, nz = small ? rz : rz / t;
    const float a = small ? 1.0f - t2 / 6.0f : sinf(t), b = small ? 0.5f - t2 / 24.0f : 1.0f - cosf(t);
    const float nn = nx * nx + ny * ny + nz * nz;
    // R = I + a S(n) + b (n n^T - (n.n) I)
    const float gRn0 = gR[0] * nx + gR[1] * ny + gR[2] * nz, gRn1 = gR[3] * nx + gR[4] * ny + gR[5] * nz,
                gRn2 = gR[6] * nx + gR[7] * ny + gR[8] * nz;                        // gR n
    const float gRtn0 = gR[0] * nx + gR[3] * ny + gR[6] * nz, gRtn1 = gR[1] * nx + gR[4] * ny + gR[7] * nz,
                gRtn2 = gR[2] * nx + gR[5] * ny + gR[8] * nz;                       // gR^T n
    float gn0 = a * s0 + b * (gRn0 + gRtn0 - 2.0f * nx * tr);
    float gn1 = a * s1 + b * (gRn1 + gRtn1 - 2.0f * ny * tr);
    float gn2 = a * s2 + b * (gRn2 + gRtn2 - 2.0f * nz * tr);
    const float ga = s0 * nx + s1 * ny + s2 * nz;                                   // <gR, S(n)>
    const float gb = (nx * gRn0 + ny * gRn1 + nz * gRn2) - nn * tr;                 // <gR, n n^T - (n.n) I>
    if (small) {
        // n = r, a = 1 - t^2/6, b = 1/2 - t^2/24 (mano.py:258-261)
        const float gt2 = -ga / 6.0f - gb / 24.0f;
        gr[0] = gn0 + 2.0f * rx * gt2; gr[1] = gn1 + 2.0f * ry * gt2; gr[2] = gn2 + 2.0f * rz * gt2;
    } else {
        // n = r / t, a = sin t, b = 1 - cos t, t = |r|
        const float gt = cosf(t) * ga + sinf(t) * gb - (gn0 * nx + gn1 * ny + gn2 * nz) / t;
        gr[0] = gn0 / t + gt * nx; gr[1] = gn1 / t + gt * ny; gr[2] = gn2 / t + gt * nz;
    }
}

__global__ void __launch_bounds__(LBS_THREADS)
lbs_bwd_kernel(const float* __restrict__ derived, const float* __restrict__ hands_mean, const float* __restrict__ rots,
               const float* __restrict__ poses, const float* __restrict__ betas, const float* __restrict__ gout,
               float* __restrict__ g_rots, float* __restrict__ g_poses, float* __restrict__ g_betas, int B) {
    pdl_sync();
    extern __shared__ __align__(16) unsigned char lbs_bwd_raw[];
    LbsBwdSmem& sm = *reinterpret_cast<LbsBwdSmem*>(lbs_bwd_raw);
    LbsSetup<BW_S>& fw = sm.fw;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = LBS_THREADS / 32;
    const int b0 = blockIdx.x * BW_S;
    const int ns = min(BW_S, B - b0);
    lbs_setup<BW_S>(fw, derived, hands_mean, rots, poses, betas, b0, ns);

    for (int e = tid; e < BW_S * (9 + 3); e += LBS_THREADS) (&sm.gRg[0][0])[e] = 0.f;    // gRg and gsum are adjacent
    for (int e = tid; e < BW_S * NJ * 12; e += LBS_THREADS) (&sm.gA[0][0][0])[e] = 0.f;
    __syncthreads();

    // ---- joint outputs 0..15 (mano.py:355-371,383-388): y_j = Rg Jtr_j - root ----
    for (int e = tid; e < BW_S * NJ; e += LBS_THREADS) {
        const int s = e / NJ, j = e % NJ;
        float gy[3] = {0.f, 0.f, 0.f};
        if (s < ns) {
            const float* g = gout + ((long long)(b0 + s) * 799 + j) * 3;
            gy[0] = g[0]; gy[1] = g[1]; gy[2] = g[2];
        }
        const float* Rg = fw.Rg[s];
#pragma unroll
        for (int c = 0; c < 3; ++c) sm.gJtr[s][j][c] = Rg[0 * 3 + c] * gy[0] + Rg[1 * 3 + c] * gy[1] + Rg[2 * 3 + c] * gy[2];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            atomicAdd(&sm.gsum[s][r], gy[r]);
#pragma unroll
            for (int c = 0; c < 3; ++c) atomicAdd(&sm.gRg[s][r * 3 + c], gy[r] * fw.Jtr[s][j][c]);
        }
    }

    // ---- phase 1: vertex sweep ----
    const float* vt_t = derived + OFF_VT;
    const float* sd_t = derived + OFF_SD;
    const float* pd_t = derived + OFF_PD;
    const float* w_t = derived + OFF_W;
    float acc_rg[BW_S][9], acc_sum[BW_S][3];
#pragma unroll
    for (int s = 0; s < BW_S; ++s) {
#pragma unroll
        for (int q = 0; q < 9; ++q) acc_rg[s][q] = 0.f;
        acc_sum[s][0] = acc_sum[s][1] = acc_sum[s][2] = 0.f;
    }
    for (int v = tid; v < NV; v += LBS_THREADS) {
        float vp[BW_S][3];
        {
            const float m0 = vt_t[v], m1 = vt_t[VP + v], m2 = vt_t[2 * VP + v];
#pragma unroll
            for (int s = 0; s < BW_S; ++s) { vp[s][0] = m0; vp[s][1] = m1; vp[s][2] = m2; }
        }
#pragma unroll 1
        for (int k = 0; k < NB + NPW; ++k) {                  // shape then pose blend shapes (mano.py:288-300)
            const float* tab = k < NB ? sd_t + (size_t)k * 3 * VP : pd_t + (size_t)(k - NB) * 3 * VP;
            const float4* wrow = reinterpret_cast<const float4*>(k < NB ? fw.betaT[k] : fw.pwT[k - NB]);
            const float d0 = tab[v], d1 = tab[VP + v], d2 = tab[2 * VP + v];
#pragma unroll
            for (int s4 = 0; s4 < BW_S / 4; ++s4) {
                const float4 w = wrow[s4];
                const float ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
                for (int u = 0; u < 4; u += 2) {              // two samples per packed FMA (same values as fmaf)
                    ffma2(vp[s4 * 4 + u][0], vp[s4 * 4 + u + 1][0], d0, ws[u], ws[u + 1]);
                    ffma2(vp[s4 * 4 + u][1], vp[s4 * 4 + u + 1][1], d1, ws[u], ws[u + 1]);
                    ffma2(vp[s4 * 4 + u][2], vp[s4 * 4 + u + 1][2], d2, ws[u], ws[u + 1]);
                }
            }
        }
        float wj[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) wj[j] = w_t[j * VP + v];
        int tip = -1;
#pragma unroll
        for (int q = 0; q < 5; ++q) if (c_tips[q] == v) tip = q;
#pragma unroll
        for (int s = 0; s < BW_S; ++s) {
            float gy[3] = {0.f, 0.f, 0.f}, x[3] = {0.f, 0.f, 0.f};
            if (s < ns) {
                const float* g = gout + ((long long)(b0 + s) * 799 + 21 + v) * 3;
                gy[0] = g[0]; gy[1] = g[1]; gy[2] = g[2];
                if (tip >= 0) {                               // the fingertip joints ARE these vertices (mano.py:373-377)
                    const float* gj = gout + ((long long)(b0 + s) * 799 + 16 + tip) * 3;
                    gy[0] += gj[0]; gy[1] += gj[1]; gy[2] += gj[2];
                }
                float T[12];
#pragma unroll
                for (int q = 0; q < 12; ++q) T[q] = 0.f;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const float4* arow = reinterpret_cast<const float4*>(fw.A[s][j]);
                    const float4 a0 = arow[0], a1 = arow[1], a2 = arow[2];
                    const float a[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
#pragma unroll
                    for (int q = 0; q < 12; q += 2) ffma2(T[q], T[q + 1], wj[j], a[q], a[q + 1]);
                }
#pragma unroll
                for (int r = 0; r < 3; ++r) x[r] = T[r * 4 + 0] * vp[s][0] + T[r * 4 + 1] * vp[s][1] + T[r * 4 + 2] * vp[s][2] + T[r * 4 + 3];
            }
            const float* Rg = fw.Rg[s];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                sm.gx[s][v][c] = Rg[0 * 3 + c] * gy[0] + Rg[1 * 3 + c] * gy[1] + Rg[2 * 3 + c] * gy[2];
                sm.vp[s][v][c] = vp[s][c];
            }
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                acc_sum[s][r] += gy[r];
#pragma unroll
                for (int c = 0; c < 3; ++c) acc_rg[s][r * 3 + c] = fmaf(gy[r], x[c], acc_rg[s][r * 3 + c]);
            }
        }
    }
#pragma unroll
    for (int s = 0; s < BW_S; ++s) {
#pragma unroll
        for (int q = 0; q < 9; ++q) {
            const float t = warp_sum(acc_rg[s][q]);
            if (lane == 0) atomicAdd(&sm.gRg[s][q], t);
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            const float t = warp_sum(acc_sum[s][r]);
            if (lane == 0) atomicAdd(&sm.gsum[s][r], t);
        }
    }
    __syncthreads();

    // ---- phase 2: gA[s][j] = sum_v w[v,j] [gx (x) vp | gx] ----
    for (int task = warp; task < BW_S * NJ; task += NW) {
        const int s = task / NJ, j = task % NJ;
        float a[12];
#pragma unroll
        for (int q = 0; q < 12; ++q) a[q] = 0.f;
        for (int v = lane; v < NV; v += 32) {
            const float w = w_t[j * VP + v];
            const float g0 = w * sm.gx[s][v][0], g1 = w * sm.gx[s][v][1], g2 = w * sm.gx[s][v][2];
            const float p0 = sm.vp[s][v][0], p1 = sm.vp[s][v][1], p2 = sm.vp[s][v][2];
            a[0] = fmaf(g0, p0, a[0]); a[1] = fmaf(g0, p1, a[1]); a[2] = fmaf(g0, p2, a[2]); a[3] += g0;
            a[4] = fmaf(g1, p0, a[4]); a[5] = fmaf(g1, p1, a[5]); a[6] = fmaf(g1, p2, a[6]); a[7] += g1;
            a[8] = fmaf(g2, p0, a[8]); a[9] = fmaf(g2, p1, a[9]); a[10] = fmaf(g2, p2, a[10]); a[11] += g2;
        }
#pragma unroll
        for (int q = 0; q < 12; ++q) {
            const float t = warp_sum(a[q]);
            if (lane == 0) sm.gA[s][j][q] = t;
        }
    }
    __syncthreads();

    // ---- phase 3: g_vposed[s][v] = T_v.R^T gx, over vp in shared memory ----
    for (int v = tid; v < NV; v += LBS_THREADS) {
        float wj[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) wj[j] = w_t[j * VP + v];
#pragma unroll 1
        for (int s = 0; s < BW_S; ++s) {
            float Tr[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) Tr[q] = 0.f;
#pragma unroll
            for (int j = 0; j < NJ; ++j) {
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int c = 0; c < 3; ++c) Tr[r * 3 + c] = fmaf(wj[j], fw.A[s][j][r * 4 + c], Tr[r * 3 + c]);
            }
            const float g0 = sm.gx[s][v][0], g1 = sm.gx[s][v][1], g2 = sm.gx[s][v][2];
#pragma unroll
            for (int c = 0; c < 3; ++c) sm.vp[s][v][c] = Tr[0 * 3 + c] * g0 + Tr[1 * 3 + c] * g1 + Tr[2 * 3 + c] * g2;
        }
    }
    __syncthreads();

    // ---- phase 4: blend-shape tables transposed: gblend[s][k] = sum_v dirs[v,:,k] . g_vposed[s][v] ----
    for (int k = warp; k < NPW + NB; k += NW) {              // k < 135: pose dirs; 135..144: shape dirs
        const float* tab = k < NPW ? pd_t + (size_t)k * 3 * VP : sd_t + (size_t)(k - NPW) * 3 * VP;
        float acc[BW_S];
#pragma unroll
        for (int s = 0; s < BW_S; ++s) acc[s] = 0.f;
        for (int v = lane; v < NV; v += 32) {
            const float d0 = tab[v], d1 = tab[VP + v], d2 = tab[2 * VP + v];
#pragma unroll
            for (int s = 0; s < BW_S; s += 2) {               // acc = d0 p0 + (d1 p1 + (d2 p2 + acc)), two samples per packed FMA
                ffma2(acc[s], acc[s + 1], d2, sm.vp[s][v][2], sm.vp[s + 1][v][2]);
                ffma2(acc[s], acc[s + 1], d1, sm.vp[s][v][1], sm.vp[s + 1][v][1]);
                ffma2(acc[s], acc[s + 1], d0, sm.vp[s][v][0], sm.vp[s + 1][v][0]);
            }
        }
#pragma unroll
        for (int s = 0; s < BW_S; ++s) {
            const float t = warp_sum(acc[s]);
            if (lane == 0) sm.gblend[s][k] = t;
        }
    }
    __syncthreads();

    // ---- phase 5: per sample, serial: chain backward, joints -> betas, Rodrigues backward ----
    if (tid < ns) {
        const int s = tid;
        const float* Rg = fw.Rg[s];
        // root = Rg Jtr_1 is subtracted from every output row: d root = -sum gy
        float gJtr[NJ][3], gGR[NJ][9], gJ[NJ][3];
#pragma unroll 1
        for (int i = 0; i < NJ; ++i) {
#pragma unroll
            for (int c = 0; c < 3; ++c) { gJtr[i][c] = sm.gJtr[s][i][c]; gJ[i][c] = 0.f; }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            gJtr[1][c] -= Rg[0 * 3 + c] * sm.gsum[s][0] + Rg[1 * 3 + c] * sm.gsum[s][1] + Rg[2 * 3 + c] * sm.gsum[s][2];
        // global rotation: y_p = Rg (p - Jtr_1)  ->  gRg = sum gy (x) p - (sum gy) (x) Jtr_1
        float gRg[9];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int c = 0; c < 3; ++c) gRg[r * 3 + c] = sm.gRg[s][r * 3 + c] - sm.gsum[s][r] * fw.Jtr[s][1][c];
        float gr[3];
        const float* rb = rots + (long long)(b0 + s) * 3;
        rodrigues_bwd(rb[0], rb[1], rb[2], gRg, gr);
        g_rots[(long long)(b0 + s) * 3 + 0] = gr[0]; g_rots[(long long)(b0 + s) * 3 + 1] = gr[1]; g_rots[(long long)(b0 + s) * 3 + 2] = gr[2];
        // A_i = [G_i.R | G_i.t - G_i.R J_i],  Jtr_i = G_i.t
        float gGt[NJ][3];
#pragma unroll 1
        for (int i = 0; i < NJ; ++i) {
            const float* gA = sm.gA[s][i];
            const float* G = fw.G[s][i];
            const float* J = fw.Jp[s][i];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                gGt[i][r] = gA[r * 4 + 3] + gJtr[i][r];
#pragma unroll
                for (int c = 0; c < 3; ++c) gGR[i][r * 3 + c] = gA[r * 4 + c] - gA[r * 4 + 3] * J[c];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) gJ[i][c] -= G[0 * 4 + c] * gA[0 * 4 + 3] + G[1 * 4 + c] * gA[1 * 4 + 3] + G[2 * 4 + c] * gA[2 * 4 + 3];
        }
        // G_i.R = G_p.R R_i,  G_i.t = G_p.R (J_i - J_p) + G_p.t   (children have larger indices than their parents)
        float* gp = g_poses + (long long)(b0 + s) * 45;
        const float* pb = poses + (long long)(b0 + s) * 45;
#pragma unroll 1
        for (int i = NJ - 1; i >= 1; --i) {
            const int p = c_parent[i];
            const float* Gp = fw.G[s][p];
            const float* Ri = fw.Rl[s][i];
            float d[3], gd[3], gRi[9];
#pragma unroll
            for (int c = 0; c < 3; ++c) d[c] = fw.Jp[s][i][c] - fw.Jp[s][p][c];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    // gG_p.R += gG_i.R R_i^T + gG_i.t (x) d ;  gR_i = G_p.R^T gG_i.R
                    gGR[p][r * 3 + c] += gGR[i][r * 3 + 0] * Ri[c * 3 + 0] + gGR[i][r * 3 + 1] * Ri[c * 3 + 1] + gGR[i][r * 3 + 2] * Ri[c * 3 + 2] +
                                         gGt[i][r] * d[c];
                    gRi[r * 3 + c] = Gp[0 * 4 + r] * gGR[i][0 * 3 + c] + Gp[1 * 4 + r] * gGR[i][1 * 3 + c] + Gp[2 * 4 + r] * gGR[i][2 * 3 + c];
                }
                gd[r] = Gp[0 * 4 + r] * gGt[i][0] + Gp[1 * 4 + r] * gGt[i][1] + Gp[2 * 4 + r] * gGt[i][2];
                gGt[p][r] += gGt[i][r];
            }
#pragma unroll
            for (int c = 0; c < 3; ++c) { gJ[i][c] += gd[c]; gJ[p][c] -= gd[c]; }
            // the pose blend shapes read R_i - I (mano.py:270-277)
#pragma unroll
            for (int q = 0; q < 9; ++q) gRi[q] += sm.gblend[s][(i - 1) * 9 + q];
            float gth[3];
            const float* hm = hands_mean + (i - 1) * 3;
            rodrigues_bwd(hm[0] + pb[(i - 1) * 3], hm[1] + pb[(i - 1) * 3 + 1], hm[2] + pb[(i - 1) * 3 + 2], gRi, gth);
            gp[(i - 1) * 3 + 0] = gth[0]; gp[(i - 1) * 3 + 1] = gth[1]; gp[(i - 1) * 3 + 2] = gth[2];
        }
        // G_0 = [R_0 | J_0] with R_0 the constant identity
#pragma unroll
        for (int c = 0; c < 3; ++c) gJ[0][c] += gGt[0][c];
        // betas: through the shape blend shapes (phase 4) and through J = J_template + J_shapedirs beta
        float* gb = g_betas + (long long)(b0 + s) * NB;
#pragma unroll 1
        for (int k = 0; k < NB; ++k) {
            float t = sm.gblend[s][NPW + k];
#pragma unroll 1
            for (int jc = 0; jc < NJ * 3; ++jc) t = fmaf(derived[OFF_JS + jc * NB + k], gJ[jc / 3][jc % 3], t);
            gb[k] = t;
        }
    }
}

}  // namespace

int launch_lbs_bwd(const float* derived, const float* hands_mean, const float* rots, const float* poses, const float* betas,
                   const float* grad_out, float* g_rots, float* g_poses, float* g_betas, int B, cudaStream_t stream) {
    SCAT_REQUIRE(derived && hands_mean && rots && poses && betas && grad_out && g_rots && g_poses && g_betas && B > 0, kErrBadArg,
                 "lbs_bwd: bad args");
    const size_t smem = sizeof(LbsBwdSmem);
    SCAT_ENSURE_SMEM(lbs_bwd_kernel, smem);
    SCAT_CHECK_CUDA(launch_k(lbs_bwd_kernel, dim3(ceil_div(B, BW_S)), dim3(LBS_THREADS), smem, stream, derived, hands_mean, rots,
                             poses, betas, grad_out, g_rots, g_poses, g_betas, B));
    SCAT_CHECK_LAUNCH();
    return 0;
}

}  // namespace scat

extern "C" int scat_lbs_bwd(const float* derived, const float* hands_mean, const float* rots, const float* poses,
                            const float* betas, const float* grad_out, float* grad_rots, float* grad_poses,
                            float* grad_betas, int32_t batch, void* stream) {
    return scat::launch_lbs_bwd(derived, hands_mean, rots, poses, betas, grad_out, grad_rots, grad_poses, grad_betas, batch,
                                (cudaStream_t)stream);
}
