// Host-side check of r2_gaussian_b200/csrc/r2x_matcalc.cuh: every gradient identity against central finite
// differences of the forward map, evaluated in double.  Built and run by tests/test_matcalc_cpu.py (nvcc, no GPU).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include "../../r2_gaussian_b200/csrc/r2x_matcalc.cuh"

using namespace r2x;
static double frand() { return (double)rand() / RAND_MAX * 2.0 - 1.0; }

// Sigma(scale, q) contracted with a symmetric weight W6 (off-diagonals weigh the single shared parameter)
static double sigma_contract(const double* sc, double mod, const double* q, const float* W6) {
    const double r = q[0], x = q[1], y = q[2], z = q[3];
    const double R[3][3] = {{1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y)},
                            {2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x)},
                            {2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)}};
    double S[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            S[i][j] = 0;
            for (int k = 0; k < 3; ++k) S[i][j] += R[i][k] * (mod * sc[k]) * (mod * sc[k]) * R[j][k];
        }
    return W6[0] * S[0][0] + W6[1] * S[0][1] + W6[2] * S[0][2] + W6[3] * S[1][1] + W6[4] * S[1][2] + W6[5] * S[2][2];
}

static double hat_contract(const double* N9, const float* V6, const float* W6) {
    const double V[3][3] = {{V6[0], V6[1], V6[2]}, {V6[1], V6[3], V6[4]}, {V6[2], V6[4], V6[5]}};
    double H[3][3];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            H[i][j] = 0;
            for (int a = 0; a < 3; ++a)
                for (int b = 0; b < 3; ++b) H[i][j] += N9[i * 3 + a] * V[a][b] * N9[j * 3 + b];
        }
    return W6[0] * H[0][0] + W6[1] * H[0][1] + W6[2] * H[0][2] + W6[3] * H[1][1] + W6[4] * H[1][2] + W6[5] * H[2][2];
}

static double inv_contract(const double* s6, const float* W6) {
    const double a = s6[0], b = s6[1], c = s6[2], d = s6[3], e = s6[4], f = s6[5];
    const double det = a * (d * f - e * e) - b * (b * f - c * e) + c * (b * e - c * d);
    const double C[6] = {(d * f - e * e) / det, (c * e - b * f) / det, (b * e - c * d) / det, (a * f - c * c) / det,
                         (b * c - a * e) / det, (a * d - b * b) / det};
    double r = 0;
    for (int k = 0; k < 6; ++k) r += W6[k] * C[k];
    return r;
}

static int fails = 0;
static void expect(const char* what, double got, double want, double scale) {
    if (fabs(got - want) > 2e-3 * scale + 1e-6) { printf("FAIL %s: got %.6g want %.6g\n", what, got, want); ++fails; }
}

int main() {
    srand(7);
    for (int trial = 0; trial < 50; ++trial) {
        // ---- cov3d_backward ----
        double sc[3] = {0.3 + 0.2 * fabs(frand()), 0.2 + 0.3 * fabs(frand()), 0.1 + 0.4 * fabs(frand())};
        double q[4] = {frand(), frand(), frand(), frand()};
        const double mod = 1.0;    // the reference's dscale convention equals the true derivative only for mod = 1
        float W6[6];
        for (int k = 0; k < 6; ++k) W6[k] = (float)frand();
        float ds[3], dr[4];
        cov3d_backward((float)sc[0], (float)sc[1], (float)sc[2], (float)mod, make_float4((float)q[0], (float)q[1], (float)q[2], (float)q[3]), W6, ds, dr);
        double gscale = 0;
        for (int k = 0; k < 3; ++k) {
            double p[3] = {sc[0], sc[1], sc[2]}, m[3] = {sc[0], sc[1], sc[2]};
            p[k] += 1e-5; m[k] -= 1e-5;
            const double fd = (sigma_contract(p, mod, q, W6) - sigma_contract(m, mod, q, W6)) / 2e-5;
            gscale = fmax(gscale, fabs(fd));
            expect("dscale", ds[k], fd, fmax(gscale, 1.0));
        }
        for (int k = 0; k < 4; ++k) {
            double p[4] = {q[0], q[1], q[2], q[3]}, m[4] = {q[0], q[1], q[2], q[3]};
            p[k] += 1e-5; m[k] -= 1e-5;
            const double fd = (sigma_contract(sc, mod, p, W6) - sigma_contract(sc, mod, m, W6)) / 2e-5;
            expect("drot", dr[k], fd, fmax(fabs(fd), 1.0));
        }
        // ---- dcov3d_from_dhat: hat = N V N^T, gradient w.r.t. V ----
        double N9[9];
        float Mm[9], V6[6];
        for (int k = 0; k < 9; ++k) { N9[k] = frand(); Mm[k] = (float)N9[k]; }
        for (int k = 0; k < 6; ++k) V6[k] = (float)frand();
        float dV[6];
        dcov3d_from_dhat(Mm, W6, dV);
        for (int k = 0; k < 6; ++k) {
            float p[6], m[6];
            for (int j = 0; j < 6; ++j) p[j] = m[j] = V6[j];
            p[k] += 1e-3f; m[k] -= 1e-3f;
            const double fd = (hat_contract(N9, p, W6) - hat_contract(N9, m, W6)) / ((double)p[k] - (double)m[k]);
            expect("dV", dV[k], fd, fmax(fabs(fd), 1.0));
        }
        // ---- dL/dN = 2 D N V ----
        {
            const Mat3 N = mat_from9(Mm);
            const Mat3 dN = matmul<false, false>(sym_grad_full(W6), matmul<false, false>(N, sym_full(V6)));
            for (int k = 0; k < 9; ++k) {
                double p[9], m[9];
                for (int j = 0; j < 9; ++j) p[j] = m[j] = N9[j];
                p[k] += 1e-5; m[k] -= 1e-5;
                const double fd = (hat_contract(p, V6, W6) - hat_contract(m, V6, W6)) / 2e-5;
                expect("dN", 2.0 * dN.m[k / 3][k % 3], fd, fmax(fabs(fd), 1.0));
            }
        }
        // ---- inverse of a symmetric positive-definite 3x3: dL/dS = -adj G adj / det^2 ----
        {
            double B[9];
            for (int k = 0; k < 9; ++k) B[k] = frand();
            double s6[6] = {0, 0, 0, 0, 0, 0};
            const int ij[6][2] = {{0, 0}, {0, 1}, {0, 2}, {1, 1}, {1, 2}, {2, 2}};
            for (int k = 0; k < 6; ++k) {
                for (int a = 0; a < 3; ++a) s6[k] += B[ij[k][0] * 3 + a] * B[ij[k][1] * 3 + a];
                if (ij[k][0] == ij[k][1]) s6[k] += 0.5;
            }
            float sf[6];
            for (int k = 0; k < 6; ++k) sf[k] = (float)s6[k];
            Mat3 K;
            const float det = sym_cofactors(sf, K);
            const Mat3 T = matmul<false, false>(K, matmul<false, false>(sym_grad_full(W6), K));
            float dh[6];
            sym_grad_pack(T, dh);
            for (int k = 0; k < 6; ++k) {
                double p[6], m[6];
                for (int j = 0; j < 6; ++j) p[j] = m[j] = s6[j];
                p[k] += 1e-6; m[k] -= 1e-6;
                const double fd = (inv_contract(p, W6) - inv_contract(m, W6)) / 2e-6;
                expect("dinv", -dh[k] / ((double)det * det), fd, fmax(fabs(fd), 1.0));
            }
        }
    }
    printf(fails ? "matcalc_check: %d FAILURES\n" : "matcalc_check: ok\n", fails);
    return fails ? 1 : 0;
}
