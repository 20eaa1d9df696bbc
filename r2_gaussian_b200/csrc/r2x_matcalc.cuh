// r2x_matcalc.cuh -- small symmetric-matrix calculus used by the per-Gaussian backward kernels.  Plain float math
// (tolerance territory), usable from host code too: tests/host/matcalc_check.cu checks every identity below against
// finite differences on the CPU.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define R2X_HD __host__ __device__ __forceinline__

namespace r2x {

// ---------------------------------------------------------------------------------------------
// Matrix calculus for the per-Gaussian chain rules (tolerance territory: plain float math).
//
// A symmetric 3x3 quantity travels as 6 numbers (00,01,02,11,12,22).  For a GRADIENT the off-diagonal numbers are
// derivatives with respect to the single parameter that fills both mirrored slots; sym_grad_full() spreads such a
// 6-vector into the full matrix (each mirrored slot gets half), after which ordinary matrix identities apply:
//     Y = N X N^T            =>  dL/dX = N^T (dL/dY) N,      dL/dN = 2 (dL/dY) N X      (X, dL/dY symmetric)
//     C = S^-1               =>  dL/dS = -C (dL/dC) C
//     Sigma = A A^T, A = R diag(s)  =>  dL/dA = 2 (dL/dSigma) A
// sym_grad_pack() folds a full gradient matrix back into the 6-vector (mirrored slots add up).
// ---------------------------------------------------------------------------------------------
struct Mat3 {
    float m[3][3];
};
R2X_HD Mat3 sym_full(const float* v6) {      // a symmetric VALUE: mirrored slots are equal
    Mat3 r;
    r.m[0][0] = v6[0]; r.m[1][1] = v6[3]; r.m[2][2] = v6[5];
    r.m[0][1] = r.m[1][0] = v6[1]; r.m[0][2] = r.m[2][0] = v6[2]; r.m[1][2] = r.m[2][1] = v6[4];
    return r;
}
R2X_HD Mat3 sym_grad_full(const float* g6) { // a symmetric GRADIENT: mirrored slots share it
    Mat3 r;
    r.m[0][0] = g6[0]; r.m[1][1] = g6[3]; r.m[2][2] = g6[5];
    r.m[0][1] = r.m[1][0] = 0.5f * g6[1]; r.m[0][2] = r.m[2][0] = 0.5f * g6[2]; r.m[1][2] = r.m[2][1] = 0.5f * g6[4];
    return r;
}
R2X_HD void sym_grad_pack(const Mat3& G, float* g6) {
    g6[0] = G.m[0][0]; g6[3] = G.m[1][1]; g6[5] = G.m[2][2];
    g6[1] = G.m[0][1] + G.m[1][0]; g6[2] = G.m[0][2] + G.m[2][0]; g6[4] = G.m[1][2] + G.m[2][1];
}
// C = op(A) op(B) with op = transpose when TA / TB
template <bool TA, bool TB>
R2X_HD Mat3 matmul(const Mat3& A, const Mat3& B) {
    Mat3 C;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 3; ++k) acc = fmaf(TA ? A.m[k][i] : A.m[i][k], TB ? B.m[j][k] : B.m[k][j], acc);
            C.m[i][j] = acc;
        }
    return C;
}
// cofactor matrix (= adjugate, the matrix is symmetric) of a symmetric 3x3 given as 6 numbers; returns det
R2X_HD float sym_cofactors(const float* h, Mat3& K) {
    K.m[0][0] = h[3] * h[5] - h[4] * h[4];
    K.m[1][1] = h[0] * h[5] - h[2] * h[2];
    K.m[2][2] = h[0] * h[3] - h[1] * h[1];
    K.m[0][1] = K.m[1][0] = h[2] * h[4] - h[1] * h[5];
    K.m[0][2] = K.m[2][0] = h[1] * h[4] - h[2] * h[3];
    K.m[1][2] = K.m[2][1] = h[1] * h[2] - h[0] * h[4];
    return h[0] * K.m[0][0] + h[1] * K.m[0][1] + h[2] * K.m[0][2];
}
// rotation matrix of the UN-normalised quaternion (r,x,y,z), textual rows (the reference never normalises here)
R2X_HD Mat3 quat_matrix(float4 q) {
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    Mat3 R;
    R.m[0][0] = 1.f - 2.f * (y * y + z * z); R.m[0][1] = 2.f * (x * y - r * z); R.m[0][2] = 2.f * (x * z + r * y);
    R.m[1][0] = 2.f * (x * y + r * z); R.m[1][1] = 1.f - 2.f * (x * x + z * z); R.m[1][2] = 2.f * (y * z - r * x);
    R.m[2][0] = 2.f * (x * z - r * y); R.m[2][1] = 2.f * (y * z + r * x); R.m[2][2] = 1.f - 2.f * (x * x + y * y);
    return R;
}

// Gradient of Sigma = A A^T, A = R(q) diag(mod * scale), w.r.t. scale (3) and the raw quaternion (4), given
// dL/dSigma as a 6-vector.  Conventions of the reference kept (RAS/backward.cu:334-397, VOX/backward.cu:180-213):
// dL/dscale is taken as if mod were 1, and R is differentiated as a function of the raw quaternion (no
// normalisation Jacobian).  dL/dq uses the split of dL/dR into its skew part (the (r, .) couplings) and its
// symmetric part (the (x,y,z) cross couplings) plus the diagonal terms.
R2X_HD void cov3d_backward(float s0, float s1, float s2, float mod, float4 q, const float* dS6,
                                               float* dscale, float* drot) {
    const Mat3 R = quat_matrix(q);
    const float s[3] = {mod * s0, mod * s1, mod * s2};
    Mat3 A;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) A.m[i][k] = R.m[i][k] * s[k];
    const Mat3 G = sym_grad_full(dS6);
    Mat3 dA = matmul<false, false>(G, A);          // dL/dA = 2 G A
    Mat3 dR;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            dA.m[i][k] *= 2.0f;
            acc = fmaf(R.m[i][k], dA.m[i][k], acc);
            dR.m[i][k] = s[k] * dA.m[i][k];
        }
        dscale[k] = acc;
    }
    const float r = q.x, x = q.y, y = q.z, z = q.w;
    const float k21 = dR.m[2][1] - dR.m[1][2], k02 = dR.m[0][2] - dR.m[2][0], k10 = dR.m[1][0] - dR.m[0][1];   // skew part
    const float p01 = dR.m[0][1] + dR.m[1][0], p02 = dR.m[0][2] + dR.m[2][0], p12 = dR.m[1][2] + dR.m[2][1];   // symmetric part
    drot[0] = 2.f * (x * k21 + y * k02 + z * k10);
    drot[1] = 2.f * (y * p01 + z * p02 + r * k21) - 4.f * x * (dR.m[1][1] + dR.m[2][2]);
    drot[2] = 2.f * (x * p01 + z * p12 + r * k02) - 4.f * y * (dR.m[0][0] + dR.m[2][2]);
    drot[3] = 2.f * (x * p02 + y * p12 + r * k10) - 4.f * z * (dR.m[0][0] + dR.m[1][1]);
}

// hat = N V N^T with N[x][y] = Mm[x*3+y]:  dL/dV (6-vector) = pack(N^T D N), D = full matrix of dL/dhat
R2X_HD Mat3 mat_from9(const float* Mm) {
    Mat3 N;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) N.m[i][j] = Mm[i * 3 + j];
    return N;
}
R2X_HD void dcov3d_from_dhat(const float* Mm, const float* dh, float* dcov) {
    const Mat3 N = mat_from9(Mm);
    const Mat3 D = sym_grad_full(dh);
    sym_grad_pack(matmul<true, false>(N, matmul<false, false>(D, N)), dcov);
}

}  // namespace r2x
