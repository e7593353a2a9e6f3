// Small dense linear algebra used by the registration hot path (all float64):
//   * symmetric 3x3 eigen-decomposition (cyclic Jacobi) -> smallest-eigenvalue eigenvector
//       replaces np.linalg.eigh in reference voxel.py:157-158 and estimate_normals.py:76-77
//   * closed-form 3x3 inverse covariance          reference voxel.py:69-102
//   * 6x6 solve with partial pivoting             reference registration.py:103 (np.linalg.solve)
//   * SO(3) exponential with the reference's small-angle branch, SE(3) right-plus
//                                                 reference math_tools.py:80-108
//   * one Gauss-Newton step incl. the stop-before-update rule   registration.py:100-111
#pragma once
#include "pcr_common.cuh"

namespace pcr {

// Eigenvector (unit) of the smallest eigenvalue of the symmetric matrix
//   [a00 a01 a02; a01 a11 a12; a02 a12 a22].
PCR_HD void smallest_eigvec_sym3(double a00, double a01, double a02, double a11, double a12, double a22,
                                 double& vx, double& vy, double& vz, double* evals /* optional [3] */) {
    double A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        double diag = fabs(A[0][0]) + fabs(A[1][1]) + fabs(A[2][2]);
        if (off <= 1e-300 || off <= 1e-17 * diag) break;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int k = 0; k < 3; ++k) {
            const int p = (k == 2) ? 1 : 0;
            const int q = (k == 0) ? 1 : 2;
            double apq = A[p][q];
            if (apq == 0.0) continue;
            double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
            double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
            double c = 1.0 / sqrt(t * t + 1.0);
            double s = t * c;
            // A <- J^T A J for the rotation in the (p,q) plane
            double app = A[p][p], aqq = A[q][q];
            A[p][p] = app - t * apq;
            A[q][q] = aqq + t * apq;
            A[p][q] = A[q][p] = 0.0;
            const int r = 3 - p - q;
            double arp = A[r][p], arq = A[r][q];
            A[r][p] = A[p][r] = c * arp - s * arq;
            A[r][q] = A[q][r] = s * arp + c * arq;
            for (int i = 0; i < 3; ++i) {
                double vip = V[i][p], viq = V[i][q];
                V[i][p] = c * vip - s * viq;
                V[i][q] = s * vip + c * viq;
            }
        }
    }
    int m = 0;
    if (A[1][1] < A[m][m]) m = 1;
    if (A[2][2] < A[m][m]) m = 2;
    vx = V[0][m]; vy = V[1][m]; vz = V[2][m];
    double nrm = sqrt(vx * vx + vy * vy + vz * vz);
    if (nrm > 0) { vx /= nrm; vy /= nrm; vz /= nrm; }
    if (evals) { evals[0] = A[0][0]; evals[1] = A[1][1]; evals[2] = A[2][2]; }
}

// Adjugate/determinant inverse, operation order of reference voxel.py:76-95 (no FMA
// contraction so that the result is bit-comparable); det == 0 -> 1e6 (quirk Q6).
// cov and icov are row-major 3x3.
PCR_HD void icov_closed_form(const double* cov, double* icov) {
    const double a = cov[0], b = cov[4], c = cov[8], d = cov[1], e = cov[2], f = cov[5];
    const double f2 = dmul_rn(f, f), d2 = dmul_rn(d, d), e2 = dmul_rn(e, e);
    const double bc = dmul_rn(b, c), ac = dmul_rn(a, c), ab = dmul_rn(a, b);
    const double dc = dmul_rn(d, c), de = dmul_rn(d, e), ef = dmul_rn(e, f);
    const double af = dmul_rn(a, f), df = dmul_rn(d, f), eb = dmul_rn(e, b);
    double det = dmul_rn(a, bc);
    det = dadd_rn(det, dmul_rn(dmul_rn(2.0, de), f));
    det = dsub_rn(det, dmul_rn(a, f2));
    det = dsub_rn(det, dmul_rn(b, e2));
    det = dsub_rn(det, dmul_rn(c, d2));
    if (det == 0.0) det = 1000000.0;
    const double c0 = ddiv_rn(dsub_rn(bc, f2), det);
    const double c1 = -ddiv_rn(dsub_rn(dc, ef), det);
    const double c2 = ddiv_rn(dsub_rn(df, eb), det);
    const double c3 = ddiv_rn(dsub_rn(ac, e2), det);
    const double c4 = -ddiv_rn(dsub_rn(af, de), det);
    const double c5 = ddiv_rn(dsub_rn(ab, d2), det);
    icov[0] = c0; icov[1] = c1; icov[2] = c2;
    icov[3] = c1; icov[4] = c3; icov[5] = c4;
    icov[6] = c2; icov[7] = c4; icov[8] = c5;
}

// Solve H x = rhs (6x6, row major, general) by LU with partial pivoting.
// Returns 0 on success, 1 if an exactly-zero pivot is met (np.linalg.solve -> LinAlgError).
// Every loop is unrolled and the row exchange is written as compare-and-swap against each candidate
// row, so that all indices are compile-time constants and the 6x7 tableau lives in registers: on the
// GPU this runs in ONE thread of the last block while the rest of the device waits (with a
// dynamically indexed tableau in local memory it was half of the accumulate kernel's duration).
PCR_HD int solve6(const double* Hin, const double* rhs, double* x) {
    double M[6][7];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; ++i) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 6; ++j) M[i][j] = Hin[i * 6 + j];
        M[i][6] = rhs[i];
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int col = 0; col < 6; ++col) {
        int piv = col;
        double best = fabs(M[col][col]);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = col + 1; r < 6; ++r) {
            const double v = fabs(M[r][col]);
            if (v > best) { best = v; piv = r; }
        }
        if (best == 0.0 || best != best) return 1;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = col + 1; r < 6; ++r) {
            if (piv == r) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int j = col; j < 7; ++j) { const double tmp = M[col][j]; M[col][j] = M[r][j]; M[r][j] = tmp; }
            }
        }
        const double inv = 1.0 / M[col][col];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int r = col + 1; r < 6; ++r) {
            const double fac = M[r][col] * inv;
            if (fac != 0.0) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
                for (int j = col + 1; j < 7; ++j) M[r][j] -= fac * M[col][j];
            }
        }
    }
    double y[6];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 5; i >= 0; --i) {
        double acc = M[i][6];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = i + 1; j < 6; ++j) acc -= M[i][j] * y[j];
        y[i] = acc / M[i][i];
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; ++i) x[i] = y[i];
    return 0;
}

// Rodrigues with the reference's first-order branch: theta^2 <= 1e-5 -> I + hat(w)
// (math_tools.py:80-98, quirk Q10: NOT orthonormal, reproduced on purpose).  R row-major.
PCR_HD void so3_exp(const double* w, double* R) {
    const double x = w[0], y = w[1], z = w[2];
    const double th2 = x * x + y * y + z * z;
    if (th2 <= 1e-5) {
        R[0] = 1;  R[1] = -z; R[2] = y;
        R[3] = z;  R[4] = 1;  R[5] = -x;
        R[6] = -y; R[7] = x;  R[8] = 1;
        return;
    }
    const double th = sqrt(th2);
    const double kx = x / th, ky = y / th, kz = z / th;
    double s, c;
    sincos(th, &s, &c);
    const double c1 = 1.0 - c;
    // K = hat(k), KK = K*K
    const double K[9] = {0, -kz, ky, kz, 0, -kx, -ky, kx, 0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double kk = K[i * 3 + 0] * K[0 * 3 + j] + K[i * 3 + 1] * K[1 * 3 + j] + K[i * 3 + 2] * K[2 * 3 + j];
            R[i * 3 + j] = (i == j ? 1.0 : 0.0) + s * K[i * 3 + j] + c1 * kk;
        }
}

// T <- T * [[Exp(dx[3:6]), dx[0:3]], [0, 1]]   (math_tools.py:101-108); T row-major 4x4.
PCR_HD void se3_plus(double* T, const double* dx) {
    double dR[9];
    so3_exp(dx + 3, dR);
    double Rn[9], tn[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j)
            Rn[i * 3 + j] = T[i * 4 + 0] * dR[0 * 3 + j] + T[i * 4 + 1] * dR[1 * 3 + j] + T[i * 4 + 2] * dR[2 * 3 + j];
        tn[i] = T[i * 4 + 0] * dx[0] + T[i * 4 + 1] * dx[1] + T[i * 4 + 2] * dx[2] + T[i * 4 + 3];
    }
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) T[i * 4 + j] = Rn[i * 3 + j];
        T[i * 4 + 3] = tn[i];
    }
}

// Expand the 29-entry record into dense H (row-major 6x6) and g.
PCR_HD void unpack_neq(const double* rec, double* H, double* g) {
    int k = 0;
    for (int i = 0; i < 6; ++i)
        for (int j = i; j < 6; ++j) { H[i * 6 + j] = rec[k]; H[j * 6 + i] = rec[k]; ++k; }
    for (int i = 0; i < 6; ++i) g[i] = rec[21 + i];
}

// One Gauss-Newton step on a reduced record (registration.py:100-111):
//   dx = -solve(H, g); if |dx| < tol: converged, T untouched (quirk Q8); else T <- T [+] dx.
// Returns 0 = updated, 1 = converged (not updated), 2 = singular H (T untouched).
PCR_HD int gauss_newton_step(const double* rec, double tol, double* T, double* dx_out, double* dx_norm_out) {
    double H[36], g[6], dx[6];
    unpack_neq(rec, H, g);
    if (solve6(H, g, dx)) return 2;
    double n2 = 0;
    for (int i = 0; i < 6; ++i) { dx[i] = -dx[i]; n2 += dx[i] * dx[i]; }
    const double nrm = sqrt(n2);
    if (dx_out) for (int i = 0; i < 6; ++i) dx_out[i] = dx[i];
    if (dx_norm_out) *dx_norm_out = nrm;
    if (nrm != nrm) return 2;
    if (nrm < tol) return 1;
    se3_plus(T, dx);
    return 0;
}

}  // namespace pcr
