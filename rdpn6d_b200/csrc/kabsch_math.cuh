// FP64 closed-form rigid alignment helpers shared by the fused solver and the batched Kabsch entry.
//
//  * kabsch3(): minimal 3-pair Kabsch.  For three pairs the centred sets are planar, so the SVD
//    solution of lib/pysixd/transform.py:940-951 (R = U diag(1,1,det) V^T) reduces to: map plane
//    normal to plane normal and rotate in-plane by atan2(sum cross, sum dot).  No iteration.
//  * rotation_from_cov(): rotation maximising tr(R^T S) for a 3x3 cross-covariance via Horn's
//    quaternion matrix N (the reference's own non-SVD branch, transform.py:953-969).  Equals the SVD
//    branch incl. the det<0 fix (:945-948).  The largest eigenvalue of N is found by Newton's method
//    on the quartic characteristic polynomial started from the upper bound (Ga+Gb)/2 (monotone
//    convergence; Theobald's QCP), its eigenvector from the adjugate of N - lambda I; a cyclic
//    Jacobi sweep on the 4x4 is the fallback when that eigenvalue is (numerically) repeated.
#pragma once
#include <cuda_runtime.h>

namespace rdpn {

#define RDPN_DEGENERATE_SIN2 1e-6

// non-degeneracy of the triangle (p0,p1,p2); explicit _rn intrinsics so that no FMA is formed and the
// result is bit-identical to oracle/pose_oracle.py:_triangle_ok (float64, one rounding per op).
__device__ __forceinline__ bool triangle_ok(const double* p0, const double* p1, const double* p2) {
    const double e1x = __dsub_rn(p1[0], p0[0]), e1y = __dsub_rn(p1[1], p0[1]), e1z = __dsub_rn(p1[2], p0[2]);
    const double e2x = __dsub_rn(p2[0], p0[0]), e2y = __dsub_rn(p2[1], p0[1]), e2z = __dsub_rn(p2[2], p0[2]);
    const double nx = __dsub_rn(__dmul_rn(e1y, e2z), __dmul_rn(e1z, e2y));
    const double ny = __dsub_rn(__dmul_rn(e1z, e2x), __dmul_rn(e1x, e2z));
    const double nz = __dsub_rn(__dmul_rn(e1x, e2y), __dmul_rn(e1y, e2x));
    const double a2 = __dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), __dmul_rn(nz, nz));
    const double l1 = __dadd_rn(__dadd_rn(__dmul_rn(e1x, e1x), __dmul_rn(e1y, e1y)), __dmul_rn(e1z, e1z));
    const double l2 = __dadd_rn(__dadd_rn(__dmul_rn(e2x, e2x), __dmul_rn(e2y, e2y)), __dmul_rn(e2z, e2z));
    return a2 > __dmul_rn(RDPN_DEGENERATE_SIN2, __dmul_rn(l1, l2));
}

__device__ __forceinline__ void plane_basis(const double* p0, const double* p1, const double* p2, double* e1,
                                            double* e2, double* n) {
    double ax = p1[0] - p0[0], ay = p1[1] - p0[1], az = p1[2] - p0[2];
    const double bx = p2[0] - p0[0], by = p2[1] - p0[1], bz = p2[2] - p0[2];
    double nx = ay * bz - az * by, ny = az * bx - ax * bz, nz = ax * by - ay * bx;
    const double il = rsqrt(ax * ax + ay * ay + az * az);  // ~1 ulp; the result is rounded to FP32 anyway
    const double in = rsqrt(nx * nx + ny * ny + nz * nz);
    ax *= il; ay *= il; az *= il;
    nx *= in; ny *= in; nz *= in;
    e1[0] = ax; e1[1] = ay; e1[2] = az;
    n[0] = nx; n[1] = ny; n[2] = nz;
    e2[0] = ny * az - nz * ay;
    e2[1] = nz * ax - nx * az;
    e2[2] = nx * ay - ny * ax;
}

// a[3][3], c[3][3]: three object / camera points (row = point).  Rt: 3x4 row-major (R | t), c ~ R a + t.
__device__ __forceinline__ void kabsch3(const double (*a)[3], const double (*c)[3], double* Rt) {
    double e1a[3], e2a[3], na[3], e1c[3], e2c[3], nc[3];
    plane_basis(a[0], a[1], a[2], e1a, e2a, na);
    plane_basis(c[0], c[1], c[2], e1c, e2c, nc);
    double ma[3], mc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        ma[k] = (a[0][k] + a[1][k] + a[2][k]) * (1.0 / 3.0);
        mc[k] = (c[0][k] + c[1][k] + c[2][k]) * (1.0 / 3.0);
    }
    double sdot = 0.0, scross = 0.0;  // m11+m22 and m21-m12 of the in-plane 2x2 covariance
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const double A0 = a[i][0] - ma[0], A1 = a[i][1] - ma[1], A2 = a[i][2] - ma[2];
        const double C0 = c[i][0] - mc[0], C1 = c[i][1] - mc[1], C2 = c[i][2] - mc[2];
        const double xa = A0 * e1a[0] + A1 * e1a[1] + A2 * e1a[2];
        const double ya = A0 * e2a[0] + A1 * e2a[1] + A2 * e2a[2];
        const double xc = C0 * e1c[0] + C1 * e1c[1] + C2 * e1c[2];
        const double yc = C0 * e2c[0] + C1 * e2c[1] + C2 * e2c[2];
        sdot += xc * xa + yc * ya;
        scross += yc * xa - xc * ya;
    }
    const double ih = rsqrt(sdot * sdot + scross * scross);
    const double cs = sdot * ih, sn = scross * ih;
    double f1[3], f2[3];  // images of e1a, e2a
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        f1[k] = cs * e1c[k] + sn * e2c[k];
        f2[k] = cs * e2c[k] - sn * e1c[k];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int k = 0; k < 3; ++k) Rt[4 * r + k] = f1[r] * e1a[k] + f2[r] * e2a[k] + nc[r] * na[k];
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) Rt[4 * r + 3] = mc[r] - (Rt[4 * r] * ma[0] + Rt[4 * r + 1] * ma[1] + Rt[4 * r + 2] * ma[2]);
}

// ---- register-lean form of kabsch3 for the fused solver ----------------------------------------------------
// One side (object or camera) of a 3-pair hypothesis: triangle test with the exact arithmetic of triangle_ok, the
// orthonormal in-plane basis (e1, n; e2 = n x e1 is recomputed where needed), the in-plane coordinates of the three
// points relative to p0 -- (0,0), (u1,0), (u2,v2) -- and the centroid.  Everything derives from the two edge vectors,
// so the 3 x 3 inputs never live as doubles.
struct TriSide {
    double e1[3], n[3];  // unit first edge, unit normal
    double u1, u2, v2;   // in-plane coordinates of p1 and p2
    double m[3];         // centroid
    bool ok;
};
__device__ __forceinline__ void tri_side(const float (*p)[3], TriSide& o) {
    const double p0x = (double)p[0][0], p0y = (double)p[0][1], p0z = (double)p[0][2];
    const double e1x = __dsub_rn((double)p[1][0], p0x), e1y = __dsub_rn((double)p[1][1], p0y), e1z = __dsub_rn((double)p[1][2], p0z);
    const double e2x = __dsub_rn((double)p[2][0], p0x), e2y = __dsub_rn((double)p[2][1], p0y), e2z = __dsub_rn((double)p[2][2], p0z);
    const double nx = __dsub_rn(__dmul_rn(e1y, e2z), __dmul_rn(e1z, e2y));
    const double ny = __dsub_rn(__dmul_rn(e1z, e2x), __dmul_rn(e1x, e2z));
    const double nz = __dsub_rn(__dmul_rn(e1x, e2y), __dmul_rn(e1y, e2x));
    const double a2 = __dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), __dmul_rn(nz, nz));
    const double l1 = __dadd_rn(__dadd_rn(__dmul_rn(e1x, e1x), __dmul_rn(e1y, e1y)), __dmul_rn(e1z, e1z));
    const double l2 = __dadd_rn(__dadd_rn(__dmul_rn(e2x, e2x), __dmul_rn(e2y, e2y)), __dmul_rn(e2z, e2z));
    o.ok = a2 > __dmul_rn(RDPN_DEGENERATE_SIN2, __dmul_rn(l1, l2));  // == triangle_ok, bit for bit
    const double il = rsqrt(l1), in = rsqrt(a2);
    o.e1[0] = e1x * il; o.e1[1] = e1y * il; o.e1[2] = e1z * il;
    o.n[0] = nx * in; o.n[1] = ny * in; o.n[2] = nz * in;
    o.u1 = l1 * il;                                   // |e1|
    o.u2 = (e1x * e2x + e1y * e2y + e1z * e2z) * il;  // e2 . e1^
    o.v2 = a2 * in * il;                              // e2 . (n^ x e1^) = |n| / |e1|
    o.m[0] = p0x + (e1x + e2x) * (1.0 / 3.0);
    o.m[1] = p0y + (e1y + e2y) * (1.0 / 3.0);
    o.m[2] = p0z + (e1z + e2z) * (1.0 / 3.0);
}
// Pose from the two sides (same closed form as kabsch3: map normal to normal, rotate in-plane by atan2(sum cross,
// sum dot) of the centred in-plane coordinates).  The object side arrives partly parked in `scr` (stride `ss`
// doubles: e1a[3], na[3], ma[0], ma[1]) so that both bases never sit in registers together.  P: 3x4 row-major FP32.
__device__ __forceinline__ void kabsch3_sides(const double* scr, int ss, double ma2, double u1a, double u2a, double v2a,
                                              const TriSide& c, float* P) {
    // centred in-plane coordinates: x = (0, u1, u2) - (u1 + u2) / 3, y = (0, 0, v2) - v2 / 3
    const double mua = (u1a + u2a) * (1.0 / 3.0), mva = v2a * (1.0 / 3.0);
    const double muc = (c.u1 + c.u2) * (1.0 / 3.0), mvc = c.v2 * (1.0 / 3.0);
    const double xa0 = -mua, xa1 = u1a - mua, xa2 = u2a - mua, ya01 = -mva, ya2 = v2a - mva;
    const double xc0 = -muc, xc1 = c.u1 - muc, xc2 = c.u2 - muc, yc01 = -mvc, yc2 = c.v2 - mvc;
    const double sdot = (xc0 * xa0 + yc01 * ya01) + (xc1 * xa1 + yc01 * ya01) + (xc2 * xa2 + yc2 * ya2);
    const double scross = (yc01 * xa0 - xc0 * ya01) + (yc01 * xa1 - xc1 * ya01) + (yc2 * xa2 - xc2 * ya2);
    const double ih = rsqrt(sdot * sdot + scross * scross);
    const double cs = sdot * ih, sn = scross * ih;
    // images of e1a, e2a: f1 = cs e1c + sn e2c, f2 = cs e2c - sn e1c with e2c = nc x e1c
    const double e2c0 = c.n[1] * c.e1[2] - c.n[2] * c.e1[1];
    const double e2c1 = c.n[2] * c.e1[0] - c.n[0] * c.e1[2];
    const double e2c2 = c.n[0] * c.e1[1] - c.n[1] * c.e1[0];
    const double f1[3] = {cs * c.e1[0] + sn * e2c0, cs * c.e1[1] + sn * e2c1, cs * c.e1[2] + sn * e2c2};
    const double f2[3] = {cs * e2c0 - sn * c.e1[0], cs * e2c1 - sn * c.e1[1], cs * e2c2 - sn * c.e1[2]};
    const double e1a0 = scr[0 * ss], e1a1 = scr[1 * ss], e1a2 = scr[2 * ss];
    const double na0 = scr[3 * ss], na1 = scr[4 * ss], na2 = scr[5 * ss];
    const double ma0 = scr[6 * ss], ma1 = scr[7 * ss];
    const double e2a0 = na1 * e1a2 - na2 * e1a1, e2a1 = na2 * e1a0 - na0 * e1a2, e2a2 = na0 * e1a1 - na1 * e1a0;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const double r0 = f1[r] * e1a0 + f2[r] * e2a0 + c.n[r] * na0;
        const double r1 = f1[r] * e1a1 + f2[r] * e2a1 + c.n[r] * na1;
        const double r2 = f1[r] * e1a2 + f2[r] * e2a2 + c.n[r] * na2;
        P[4 * r + 0] = (float)r0;
        P[4 * r + 1] = (float)r1;
        P[4 * r + 2] = (float)r2;
        P[4 * r + 3] = (float)(c.m[r] - (r0 * ma0 + r1 * ma1 + r2 * ma2));
    }
}

// S[i*3+j] = sum_w c_i a_j (camera row, object column), i.e. v1 . v0^T of transform.py:942.
// R (row-major) maximises sum c^T R a over SO(3).
__device__ __forceinline__ void quat_to_rot(double q0, double q1, double q2, double q3, double* R) {
    const double inv = 1.0 / sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3);
    q0 *= inv; q1 *= inv; q2 *= inv; q3 *= inv;
    R[0] = 1.0 - 2.0 * (q2 * q2 + q3 * q3); R[1] = 2.0 * (q1 * q2 - q3 * q0);       R[2] = 2.0 * (q1 * q3 + q2 * q0);
    R[3] = 2.0 * (q1 * q2 + q3 * q0);       R[4] = 1.0 - 2.0 * (q1 * q1 + q3 * q3); R[5] = 2.0 * (q2 * q3 - q1 * q0);
    R[6] = 2.0 * (q1 * q3 - q2 * q0);       R[7] = 2.0 * (q2 * q3 + q1 * q0);       R[8] = 1.0 - 2.0 * (q1 * q1 + q2 * q2);
}

__device__ __forceinline__ double det3(double a, double b, double c, double d, double e, double f, double g, double h,
                                       double i) {
    return a * (e * i - f * h) - b * (d * i - f * g) + c * (d * h - e * g);
}

static __device__ __noinline__ void rotation_from_cov_jacobi(const double* S, double* R);

// Ga = sum_w |a - a_mean|^2, Gb = sum_w |c - c_mean|^2 (only used as the Newton start: any upper bound of
// the largest eigenvalue works).
__device__ __forceinline__ void rotation_from_cov(const double* S, double Ga, double Gb, double* R) {
    const double Sxx = S[0], Sxy = S[3], Sxz = S[6];
    const double Syx = S[1], Syy = S[4], Syz = S[7];
    const double Szx = S[2], Szy = S[5], Szz = S[8];
    // symmetric traceless N (upper triangle)
    const double n00 = Sxx + Syy + Szz, n01 = Syz - Szy, n02 = Szx - Sxz, n03 = Sxy - Syx;
    const double n11 = Sxx - Syy - Szz, n12 = Sxy + Syx, n13 = Szx + Sxz;
    const double n22 = -Sxx + Syy - Szz, n23 = Syz + Szy;
    const double n33 = -Sxx - Syy + Szz;
    // characteristic polynomial  l^4 + C2 l^2 + C1 l + C0
    double ss = 0.0;
#pragma unroll
    for (int i = 0; i < 9; ++i) ss += S[i] * S[i];
    const double C2 = -2.0 * ss;
    const double C1 = -8.0 * det3(S[0], S[1], S[2], S[3], S[4], S[5], S[6], S[7], S[8]);
    const double C0 = n00 * det3(n11, n12, n13, n12, n22, n23, n13, n23, n33) -
                      n01 * det3(n01, n12, n13, n02, n22, n23, n03, n23, n33) +
                      n02 * det3(n01, n11, n13, n02, n12, n23, n03, n13, n33) -
                      n03 * det3(n01, n11, n12, n02, n12, n22, n03, n13, n23);
    double lam = 0.5 * (Ga + Gb);
    const double scale = lam;
    bool converged = false;
    for (int it = 0; it < 64; ++it) {
        const double x2 = lam * lam;
        const double b = (x2 + C2) * lam;
        const double a = b + C1;
        const double fp = 2.0 * x2 * lam + b + a;
        if (fp == 0.0) break;
        const double delta = (a * lam + C0) / fp;
        lam -= delta;
        if (fabs(delta) <= 1e-15 * fabs(lam)) { converged = true; break; }
    }
    // eigenvector: column of adj(N - lam I) with the largest diagonal cofactor
    const double m00 = n00 - lam, m11 = n11 - lam, m22 = n22 - lam, m33 = n33 - lam;
    const double c00 = det3(m11, n12, n13, n12, m22, n23, n13, n23, m33);
    const double c11 = det3(m00, n02, n03, n02, m22, n23, n03, n23, m33);
    const double c22 = det3(m00, n01, n03, n01, m11, n13, n03, n13, m33);
    const double c33 = det3(m00, n01, n02, n01, m11, n12, n02, n12, m22);
    const double a00 = fabs(c00), a11 = fabs(c11), a22 = fabs(c22), a33 = fabs(c33);
    const double big = fmax(fmax(a00, a11), fmax(a22, a33));
    // |cofactor| ~ product of the three eigenvalue gaps: tiny relative to scale^3 => repeated eigenvalue
    if (!converged || !(big > 1e-18 * scale * scale * scale)) {
        rotation_from_cov_jacobi(S, R);
        return;
    }
    double q0, q1, q2, q3;
    if (big == a00) {
        q0 = c00;
        q1 = -det3(n01, n12, n13, n02, m22, n23, n03, n23, m33);
        q2 = det3(n01, m11, n13, n02, n12, n23, n03, n13, m33);
        q3 = -det3(n01, m11, n12, n02, n12, m22, n03, n13, n23);
    } else if (big == a11) {
        q0 = -det3(n01, n02, n03, n12, m22, n23, n13, n23, m33);
        q1 = c11;
        q2 = -det3(m00, n02, n03, n01, n12, n13, n03, n23, m33);
        q3 = det3(m00, n02, n03, n01, n12, n13, n02, m22, n23);
    } else if (big == a22) {
        q0 = det3(n01, n02, n03, m11, n12, n13, n13, n23, m33);
        q1 = -det3(m00, n02, n03, n01, n12, n13, n03, n23, m33);
        q2 = c22;
        q3 = -det3(m00, n01, n03, n01, m11, n13, n02, n12, n23);
    } else {
        q0 = -det3(n01, n02, n03, m11, n12, n13, n12, m22, n23);
        q1 = det3(m00, n02, n03, n01, n12, n13, n02, m22, n23);
        q2 = -det3(m00, n01, n03, n01, m11, n13, n02, n12, n23);
        q3 = c33;
    }
    quat_to_rot(q0, q1, q2, q3, R);
}

static __device__ __noinline__ void rotation_from_cov_jacobi(const double* S, double* R) {
    // Horn's N with Sh_ij = sum a_i c_j = S[j*3+i]
    const double Sxx = S[0], Sxy = S[3], Sxz = S[6];
    const double Syx = S[1], Syy = S[4], Syz = S[7];
    const double Szx = S[2], Szy = S[5], Szz = S[8];
    double A[4][4], V[4][4];
    A[0][0] = Sxx + Syy + Szz; A[0][1] = Syz - Szy;        A[0][2] = Szx - Sxz;         A[0][3] = Sxy - Syx;
    A[1][1] = Sxx - Syy - Szz; A[1][2] = Sxy + Syx;        A[1][3] = Szx + Sxz;
    A[2][2] = -Sxx + Syy - Szz; A[2][3] = Syz + Szy;
    A[3][3] = -Sxx - Syy + Szz;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < i) A[i][j] = A[j][i];
            V[i][j] = (i == j) ? 1.0 : 0.0;
        }
    for (int sweep = 0; sweep < 24; ++sweep) {
        double off = 0.0, diag = 0.0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            diag += A[i][i] * A[i][i];
#pragma unroll
            for (int j = i + 1; j < 4; ++j) off += A[i][j] * A[i][j];
        }
        if (off <= 1e-32 * diag || off == 0.0) break;
#pragma unroll
        for (int p = 0; p < 3; ++p)
#pragma unroll
            for (int q = p + 1; q < 4; ++q) {
                const double apq = A[p][q];
                if (apq != 0.0) {
                    const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
                    const double tt = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                    const double cc = 1.0 / sqrt(tt * tt + 1.0), ss = tt * cc;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // A <- A J
                        const double akp = A[k][p], akq = A[k][q];
                        A[k][p] = cc * akp - ss * akq;
                        A[k][q] = ss * akp + cc * akq;
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k) {  // A <- J^T A ; V <- V J
                        const double apk = A[p][k], aqk = A[q][k];
                        A[p][k] = cc * apk - ss * aqk;
                        A[q][k] = ss * apk + cc * aqk;
                        const double vkp = V[k][p], vkq = V[k][q];
                        V[k][p] = cc * vkp - ss * vkq;
                        V[k][q] = ss * vkp + cc * vkq;
                    }
                }
            }
    }
    int best = 0;
    double bv = A[0][0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
        if (A[i][i] > bv) { bv = A[i][i]; best = i; }
    double q0 = 0, q1 = 0, q2 = 0, q3 = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (i == best) { q0 = V[0][i]; q1 = V[1][i]; q2 = V[2][i]; q3 = V[3][i]; }
    quat_to_rot(q0, q1, q2, q3, R);
}

}  // namespace rdpn
