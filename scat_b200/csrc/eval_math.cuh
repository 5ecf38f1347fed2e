// Per-sample arithmetic of the evaluation metrics, written once for device and host:
// the kernels in eval_metrics.cu call it per thread, and tests/host/eval_math_host.cpp compiles the SAME functions with
// g++ so the CPU test suite checks the arithmetic against the oracle without a GPU.
//
//   similarity_align  <- batch_compute_similarity_transform_torch, /root/reference/eval.py:110-161
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define SCAT_HD __host__ __device__ __forceinline__
#else
#define SCAT_HD inline
#endif

namespace scat {
namespace evalm {

// eigen-decomposition of a symmetric 3x3 (cyclic Jacobi, double): a = v diag(w) v^T, columns of v are eigenvectors,
// eigenvalues sorted descending
SCAT_HD void jacobi_eig3(double a[3][3], double w[3], double v[3][3]) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) v[i][j] = i == j ? 1.0 : 0.0;
    for (int sweep = 0; sweep < 12; ++sweep) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        const double diag = fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]);
        if (off <= 1e-30 || off <= 1e-17 * diag) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                if (fabs(a[p][q]) <= 1e-300) continue;
                const double theta = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {                 // a <- a J   (columns p, q)
                    const double akp = a[k][p], akq = a[k][q];
                    a[k][p] = c * akp - s * akq;
                    a[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {                 // a <- J^T a (rows p, q)
                    const double apk = a[p][k], aqk = a[q][k];
                    a[p][k] = c * apk - s * aqk;
                    a[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    const double vkp = v[k][p], vkq = v[k][q];
                    v[k][p] = c * vkp - s * vkq;
                    v[k][q] = s * vkp + c * vkq;
                }
            }
    }
    for (int i = 0; i < 3; ++i) w[i] = a[i][i];
    for (int i = 0; i < 2; ++i)                               // sort descending (3 elements)
        for (int j = 0; j < 2 - i; ++j)
            if (w[j] < w[j + 1]) {
                const double tw = w[j]; w[j] = w[j + 1]; w[j + 1] = tw;
                for (int k = 0; k < 3; ++k) { const double tv = v[k][j]; v[k][j] = v[k][j + 1]; v[k][j + 1] = tv; }
            }
}

SCAT_HD double det3(const double m[3][3]) {
    return m[0][0] * (m[1][1] * m[2][2] - m[1][2] * m[2][1]) - m[0][1] * (m[1][0] * m[2][2] - m[1][2] * m[2][0]) +
           m[0][2] * (m[1][0] * m[2][1] - m[1][1] * m[2][0]);
}

// Similarity transform (s, R, t) that takes the n points `s1` closest to `s2` (both [n,3], fp32, row stride 3), and
// the transformed points `out` = s R s1 + t.  Steps as eval.py:124-156: remove means; var1 = sum |X1|^2;
// K = X1^T-outer-X2 (3x3); K = U S V^T; Z = diag(1, 1, sign det(U V^T)); R = V Z U^T; s = tr(R K) / var1;
// t = mu2 - s R mu1.  The SVD comes from the eigen-decomposition of K^T K (V, S^2) and U = K V S^-1.
SCAT_HD void similarity_align(const float* s1, const float* s2, int n, float* out, float* scale_out) {
    double mu1[3] = {0, 0, 0}, mu2[3] = {0, 0, 0};
    for (int j = 0; j < n; ++j)
        for (int c = 0; c < 3; ++c) { mu1[c] += s1[j * 3 + c]; mu2[c] += s2[j * 3 + c]; }
    for (int c = 0; c < 3; ++c) { mu1[c] /= n; mu2[c] /= n; }
    double K[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, var1 = 0;
    for (int j = 0; j < n; ++j) {
        double x1[3], x2[3];
        for (int c = 0; c < 3; ++c) { x1[c] = s1[j * 3 + c] - mu1[c]; x2[c] = s2[j * 3 + c] - mu2[c]; }
        for (int a = 0; a < 3; ++a) {
            var1 += x1[a] * x1[a];
            for (int b = 0; b < 3; ++b) K[a][b] += x1[a] * x2[b];
        }
    }
    double B[3][3], w[3], V[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) B[a][b] = K[0][a] * K[0][b] + K[1][a] * K[1][b] + K[2][a] * K[2][b];
    jacobi_eig3(B, w, V);
    // u1, u2 = K v_i / s_i (u2 re-orthogonalised); u3 = +-(u1 x u2), the sign taken from K v3.  When the third singular
    // value vanishes (coplanar points, or n = 3) that sign is noise, and R below does not depend on it: flipping u3 flips
    // det(U V^T) with it.
    double U[3][3], Kv[3][3];
    for (int i = 0; i < 3; ++i)
        for (int a = 0; a < 3; ++a) Kv[a][i] = K[a][0] * V[0][i] + K[a][1] * V[1][i] + K[a][2] * V[2][i];
    double n1 = sqrt(Kv[0][0] * Kv[0][0] + Kv[1][0] * Kv[1][0] + Kv[2][0] * Kv[2][0]);
    n1 = n1 > 0 ? 1.0 / n1 : 0.0;
    for (int a = 0; a < 3; ++a) U[a][0] = Kv[a][0] * n1;
    const double d12 = U[0][0] * Kv[0][1] + U[1][0] * Kv[1][1] + U[2][0] * Kv[2][1];
    double u2[3] = {Kv[0][1] - d12 * U[0][0], Kv[1][1] - d12 * U[1][0], Kv[2][1] - d12 * U[2][0]};
    double n2 = sqrt(u2[0] * u2[0] + u2[1] * u2[1] + u2[2] * u2[2]);
    n2 = n2 > 0 ? 1.0 / n2 : 0.0;
    for (int a = 0; a < 3; ++a) U[a][1] = u2[a] * n2;
    U[0][2] = U[1][0] * U[2][1] - U[2][0] * U[1][1];
    U[1][2] = U[2][0] * U[0][1] - U[0][0] * U[2][1];
    U[2][2] = U[0][0] * U[1][1] - U[1][0] * U[0][1];
    if (U[0][2] * Kv[0][2] + U[1][2] * Kv[1][2] + U[2][2] * Kv[2][2] < 0)
        for (int a = 0; a < 3; ++a) U[a][2] = -U[a][2];
    double UVt[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) UVt[a][b] = U[a][0] * V[b][0] + U[a][1] * V[b][1] + U[a][2] * V[b][2];
    const double d = det3(UVt);
    const double z = d > 0 ? 1.0 : (d < 0 ? -1.0 : 0.0);
    double R[3][3];
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) R[a][b] = V[a][0] * U[b][0] + V[a][1] * U[b][1] + z * V[a][2] * U[b][2];
    double tr = 0;
    for (int a = 0; a < 3; ++a)
        for (int b = 0; b < 3; ++b) tr += R[a][b] * K[b][a];
    const double s = tr / var1;
    double t[3];
    for (int a = 0; a < 3; ++a) t[a] = mu2[a] - s * (R[a][0] * mu1[0] + R[a][1] * mu1[1] + R[a][2] * mu1[2]);
    for (int j = 0; j < n; ++j) {                            // `out` may alias `s1`
        const double x = s1[j * 3], y = s1[j * 3 + 1], zc = s1[j * 3 + 2];
        for (int a = 0; a < 3; ++a) out[j * 3 + a] = (float)(s * (R[a][0] * x + R[a][1] * y + R[a][2] * zc) + t[a]);
    }
    if (scale_out) *scale_out = (float)s;
}

}  // namespace evalm
}  // namespace scat
