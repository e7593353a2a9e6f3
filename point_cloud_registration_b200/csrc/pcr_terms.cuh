// Per-correspondence residual / Jacobian terms of the four registration variants and the
// final assembly of the 29-entry normal-equation record.  float32 per-point geometry (the
// reference transforms and searches in float32, quirk Q7), accumulated by the caller.
//
//   ICP     reference icp.py:40-56                  (incl. quirk Q1: g_rot = sum p x (R r))
//   PLANE   reference plane_icp.py:46-67
//   VPLANE  reference voxelized_plane_icp.py:41-62  (same algebra as PLANE)
//   NDT     reference ndt.py:39-56
#pragma once
#include "pcr_common.cuh"

namespace pcr {

struct Pose32 {          // float32 cast of the current 4x4 transform (row-major R, t)
    float r[9];
    float t[3];
};

PCR_HD void pose32_from_T(const double* T, Pose32& P) {
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) P.r[i * 3 + j] = (float)T[i * 4 + j];
        P.t[i] = (float)T[i * 4 + 3];
    }
}

// src = R p + t in float32 (math_tools.py:111-113 with T cast to float32 by the callers)
PCR_HD void transform32(const Pose32& P, float px, float py, float pz, float& sx, float& sy, float& sz) {
    sx = fmaf(P.r[2], pz, fmaf(P.r[1], py, P.r[0] * px)) + P.t[0];
    sy = fmaf(P.r[5], pz, fmaf(P.r[4], py, P.r[3] * px)) + P.t[1];
    sz = fmaf(P.r[8], pz, fmaf(P.r[7], py, P.r[6] * px)) + P.t[2];
}

// ---- ICP: raw sums, assembled later by assemble_icp --------------------------------------
// acc: [0] count, [1..3] sum p, [4..9] sum (xx,yy,zz,xy,xz,yz) of p, [10..12] sum r,
//      [13..15] sum p x (R r), [16] sum r.r
template <typename A>
PCR_HD void accum_icp(A* acc, const Pose32& P, float px, float py, float pz, float rx, float ry, float rz) {
    // (operands are cast to the accumulator type BEFORE the product: with float64 accumulators every term is
    //  one DFMA on converted operands instead of a float product, a conversion and an add)
    const A X = (A)px, Y = (A)py, Z = (A)pz;
    acc[0] += (A)1;
    acc[1] += X; acc[2] += Y; acc[3] += Z;
    acc[4] += X * X; acc[5] += Y * Y; acc[6] += Z * Z;
    acc[7] += X * Y; acc[8] += X * Z; acc[9] += Y * Z;
    acc[10] += (A)rx; acc[11] += (A)ry; acc[12] += (A)rz;
    const float vx = P.r[0] * rx + P.r[1] * ry + P.r[2] * rz;      // v = R r   (quirk Q1)
    const float vy = P.r[3] * rx + P.r[4] * ry + P.r[5] * rz;
    const float vz = P.r[6] * rx + P.r[7] * ry + P.r[8] * rz;
    acc[13] += Y * (A)vz - Z * (A)vy;
    acc[14] += Z * (A)vx - X * (A)vz;
    acc[15] += X * (A)vy - Y * (A)vx;
    acc[16] += (A)rx * (A)rx + (A)ry * (A)ry + (A)rz * (A)rz;
}

// raw[17] (reduced, float64) + current T -> rec[29]
PCR_HD void assemble_icp(const double* raw, const double* T, double* rec) {
    const double n = raw[0], sx = raw[1], sy = raw[2], sz = raw[3];
    const double xx = raw[4], yy = raw[5], zz = raw[6], xy = raw[7], xz = raw[8], yz = raw[9];
    // H_tr = -R * hat(sum p)
    const double K[9] = {0, -sz, sy, sz, 0, -sx, -sy, sx, 0};
    double Htr[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            Htr[i * 3 + j] = -(T[i * 4 + 0] * K[0 * 3 + j] + T[i * 4 + 1] * K[1 * 3 + j] + T[i * 4 + 2] * K[2 * 3 + j]);
    rec[0] = n;  rec[1] = 0;  rec[2] = 0;  rec[3] = Htr[0]; rec[4] = Htr[1]; rec[5] = Htr[2];
    rec[6] = n;  rec[7] = 0;  rec[8] = Htr[3]; rec[9] = Htr[4]; rec[10] = Htr[5];
    rec[11] = n; rec[12] = Htr[6]; rec[13] = Htr[7]; rec[14] = Htr[8];
    rec[15] = zz + yy; rec[16] = -xy; rec[17] = -xz;
    rec[18] = xx + zz; rec[19] = -yz;
    rec[20] = xx + yy;
    rec[21] = raw[10]; rec[22] = raw[11]; rec[23] = raw[12];
    rec[24] = raw[13]; rec[25] = raw[14]; rec[26] = raw[15];
    rec[27] = raw[16];
    rec[28] = n;
}

// ---- PLANE / VPLANE: r = n.(src - q), J = [n^T, (p x R^T n)^T]; acc is already rec-shaped ---
template <typename A>
PCR_HD void accum_plane(A* acc, const Pose32& P, float px, float py, float pz,
                        float dx, float dy, float dz, float nx, float ny, float nz) {
    const float r = nx * dx + ny * dy + nz * dz;
    const float ax = P.r[0] * nx + P.r[3] * ny + P.r[6] * nz;      // a = R^T n
    const float ay = P.r[1] * nx + P.r[4] * ny + P.r[7] * nz;
    const float az = P.r[2] * nx + P.r[5] * ny + P.r[8] * nz;
    float J[6];
    J[0] = nx; J[1] = ny; J[2] = nz;
    J[3] = py * az - pz * ay;
    J[4] = pz * ax - px * az;
    J[5] = px * ay - py * ax;
    A Ja[6];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; ++i) Ja[i] = (A)J[i];
    const A ra = (A)r;
    int k = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; ++i) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = i; j < 6; ++j) acc[k++] += Ja[i] * Ja[j];
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 6; ++i) acc[21 + i] += Ja[i] * ra;
    acc[27] += ra * ra;
    acc[28] += (A)1;
}

// ---- NDT: d = src - mu, J = [I, -R hat(p)], weight W = Sigma^-1 (symmetric, 6 unique) -------
// w6 = (W00, W01, W02, W11, W12, W22); acc is rec-shaped.
template <typename A>
PCR_HD void accum_ndt(A* acc, const Pose32& P, float px, float py, float pz,
                      float dx, float dy, float dz, const float* w6) {
    const float W[9] = {w6[0], w6[1], w6[2], w6[1], w6[3], w6[4], w6[2], w6[4], w6[5]};
    const float* R = P.r;
    // q = W d ; e2 = d.q ; g_t = q ; g_r = p x (R^T q)
    const float qx = W[0] * dx + W[1] * dy + W[2] * dz;
    const float qy = W[3] * dx + W[4] * dy + W[5] * dz;
    const float qz = W[6] * dx + W[7] * dy + W[8] * dz;
    const float bx = R[0] * qx + R[3] * qy + R[6] * qz;
    const float by = R[1] * qx + R[4] * qy + R[7] * qz;
    const float bz = R[2] * qx + R[5] * qy + R[8] * qz;
    acc[21] += qx; acc[22] += qy; acc[23] += qz;
    acc[24] += py * bz - pz * by;
    acc[25] += pz * bx - px * bz;
    acc[26] += px * by - py * bx;
    acc[27] += dx * qx + dy * qy + dz * qz;
    acc[28] += 1.0f;
    // H_tt = W
    acc[0] += W[0]; acc[1] += W[1]; acc[2] += W[2]; acc[6] += W[4]; acc[7] += W[5]; acc[11] += W[8];
    // C = W R ; H_tr row i = p x C_i
    float C[9];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 3; ++i) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int j = 0; j < 3; ++j) C[i * 3 + j] = W[i * 3 + 0] * R[0 * 3 + j] + W[i * 3 + 1] * R[1 * 3 + j] + W[i * 3 + 2] * R[2 * 3 + j];
    }
    acc[3] += py * C[2] - pz * C[1];  acc[4] += pz * C[0] - px * C[2];  acc[5] += px * C[1] - py * C[0];
    acc[8] += py * C[5] - pz * C[4];  acc[9] += pz * C[3] - px * C[5];  acc[10] += px * C[4] - py * C[3];
    acc[12] += py * C[8] - pz * C[7]; acc[13] += pz * C[6] - px * C[8]; acc[14] += px * C[7] - py * C[6];
    // M = R^T C (symmetric) ; H_rr = U^T M U with U = hat(p), columns u0=(0,pz,-py) u1=(-pz,0,px) u2=(py,-px,0)
    const float m00 = R[0] * C[0] + R[3] * C[3] + R[6] * C[6];
    const float m01 = R[0] * C[1] + R[3] * C[4] + R[6] * C[7];
    const float m02 = R[0] * C[2] + R[3] * C[5] + R[6] * C[8];
    const float m11 = R[1] * C[1] + R[4] * C[4] + R[7] * C[7];
    const float m12 = R[1] * C[2] + R[4] * C[5] + R[7] * C[8];
    const float m22 = R[2] * C[2] + R[5] * C[5] + R[8] * C[8];
    // N_j = M u_j
    const float n0y = m11 * pz - m12 * py, n0z = m12 * pz - m22 * py;   // (N_0.x is never needed: u_0.x = 0)
    const float n1x = -m00 * pz + m02 * px, n1y = -m01 * pz + m12 * px, n1z = -m02 * pz + m22 * px;
    const float n2x = m00 * py - m01 * px, n2y = m01 * py - m11 * px, n2z = m02 * py - m12 * px;
    // H_rr[i][j] = u_i . N_j
    acc[15] += pz * n0y - py * n0z;          // u0.N0
    acc[16] += pz * n1y - py * n1z;          // u0.N1
    acc[17] += pz * n2y - py * n2z;          // u0.N2
    acc[18] += -pz * n1x + px * n1z;         // u1.N1
    acc[19] += -pz * n2x + px * n2z;         // u1.N2
    acc[20] += py * n2x - px * n2y;          // u2.N2
}

}  // namespace pcr
