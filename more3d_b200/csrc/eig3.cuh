// Symmetric 3x3 eigen-solve shared by the PCA-normal kernels.
#pragma once
#include "common.cuh"

namespace m3d {

// ------------------------------------------------------------------------------------------
// symmetric 3x3 eigen-solve: smallest eigenpair, closed form with a Jacobi fallback
// ------------------------------------------------------------------------------------------
static __device__ __noinline__ void jacobi3_min(double a00, double a01, double a02, double a11, double a12,
                                         double a22, double& lmin, double v[3]) {
    double A[3][3] = {{a00, a01, a02}, {a01, a11, a12}, {a02, a12, a22}};
    double V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    for (int sweep = 0; sweep < 12; ++sweep) {
        double off = fabs(A[0][1]) + fabs(A[0][2]) + fabs(A[1][2]);
        if (off < 1e-300) break;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                double apq = A[p][q];
                if (fabs(apq) < 1e-300) continue;
                double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                double t = ((theta >= 0) ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < 3; ++k) {
                    double akp = A[k][p], akq = A[k][q];
                    A[k][p] = c * akp - s * akq;
                    A[k][q] = s * akp + c * akq;
                }
                for (int k = 0; k < 3; ++k) {
                    double apk = A[p][k], aqk = A[q][k];
                    A[p][k] = c * apk - s * aqk;
                    A[q][k] = s * apk + c * aqk;
                }
                for (int k = 0; k < 3; ++k) {
                    double vkp = V[k][p], vkq = V[k][q];
                    V[k][p] = c * vkp - s * vkq;
                    V[k][q] = s * vkp + c * vkq;
                }
            }
    }
    int m = 0;
    if (A[1][1] < A[m][m]) m = 1;
    if (A[2][2] < A[m][m]) m = 2;
    lmin = A[m][m];
    v[0] = V[0][m]; v[1] = V[1][m]; v[2] = V[2][m];
}

// Input: symmetric matrix (any scale).  Output: smallest eigenvalue (same scale) and unit eigenvector.
__device__ __forceinline__ void sym3_min_eig(double a00, double a01, double a02, double a11, double a12,
                                             double a22, double& lmin, double v[3]) {
    double sc = fmax(fmax(fabs(a00), fabs(a11)), fmax(fabs(a22), fmax(fabs(a01), fmax(fabs(a02), fabs(a12)))));
    if (!(sc > 0.0)) { lmin = 0.0; v[0] = 1.0; v[1] = 0.0; v[2] = 0.0; return; }
    const double is = 1.0 / sc;
    a00 *= is; a01 *= is; a02 *= is; a11 *= is; a12 *= is; a22 *= is;
    const double p1 = a01 * a01 + a02 * a02 + a12 * a12;
    const double q = (a00 + a11 + a22) * (1.0 / 3.0);
    const double b00 = a00 - q, b11 = a11 - q, b22 = a22 - q;
    const double p2 = b00 * b00 + b11 * b11 + b22 * b22 + 2.0 * p1;
    bool ok = false;
    double lam = 0.0;
    if (p2 > 0.0) {
        // Smith's trigonometric solution of the characteristic cubic
        const double p = sqrt(p2 * (1.0 / 6.0));
        const double ip = 1.0 / p;
        const double c00 = b00 * ip, c11 = b11 * ip, c22 = b22 * ip;
        const double c01 = a01 * ip, c02 = a02 * ip, c12 = a12 * ip;
        double hd = 0.5 * (c00 * (c11 * c22 - c12 * c12) - c01 * (c01 * c22 - c12 * c02) +
                           c02 * (c01 * c12 - c11 * c02));
        hd = fmin(fmax(hd, -1.0), 1.0);
        const double phi = acos(hd) * (1.0 / 3.0);
        lam = q + 2.0 * p * cos(phi + 2.0943951023931954923);  // + 2*pi/3 -> smallest root
        // eigenvector = largest cross product of two rows of (A - lam I)
        const double r0x = a00 - lam, r0y = a01, r0z = a02;
        const double r1x = a01, r1y = a11 - lam, r1z = a12;
        const double r2x = a02, r2y = a12, r2z = a22 - lam;
        double cx0 = r0y * r1z - r0z * r1y, cy0 = r0z * r1x - r0x * r1z, cz0 = r0x * r1y - r0y * r1x;
        double cx1 = r0y * r2z - r0z * r2y, cy1 = r0z * r2x - r0x * r2z, cz1 = r0x * r2y - r0y * r2x;
        double cx2 = r1y * r2z - r1z * r2y, cy2 = r1z * r2x - r1x * r2z, cz2 = r1x * r2y - r1y * r2x;
        double n0 = cx0 * cx0 + cy0 * cy0 + cz0 * cz0;
        double n1 = cx1 * cx1 + cy1 * cy1 + cz1 * cz1;
        double n2 = cx2 * cx2 + cy2 * cy2 + cz2 * cz2;
        double bx = cx0, by = cy0, bz = cz0, bn = n0;
        if (n1 > bn) { bx = cx1; by = cy1; bz = cz1; bn = n1; }
        if (n2 > bn) { bx = cx2; by = cy2; bz = cz2; bn = n2; }
        if (bn > 1e-22) {
            const double in = 1.0 / sqrt(bn);
            bx *= in; by *= in; bz *= in;
            // Rayleigh quotient + residual check
            const double wx = a00 * bx + a01 * by + a02 * bz;
            const double wy = a01 * bx + a11 * by + a12 * bz;
            const double wz = a02 * bx + a12 * by + a22 * bz;
            const double rq = bx * wx + by * wy + bz * wz;
            const double ex = wx - rq * bx, ey = wy - rq * by, ez = wz - rq * bz;
            if (ex * ex + ey * ey + ez * ez < 1e-24) {
                lam = rq; v[0] = bx; v[1] = by; v[2] = bz; ok = true;
            }
        }
    }
    if (!ok) jacobi3_min(a00, a01, a02, a11, a12, a22, lam, v);
    lmin = lam * sc;
}


}  // namespace m3d
