/*
 * oracle/pose_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * The float32 inner loops of the composite pose-solve oracle (oracle/pose_oracle.py), in plain
 * C so that fused multiply-adds can be stated explicitly (numpy has no fma).  What they restate:
 *
 *   rigid apply  R*p + t            /root/reference/lib/pysixd/misc.py:895-905 (transform_pts_Rt),
 *                                   lib/pysixd/pose_error.py:264-274
 *   residual of ALL points, L2 norm /root/reference/lib/pysixd/misc.py:108-109
 *   inliers = errs < thr (strict)   /root/reference/lib/pysixd/misc.py:111
 *
 * Arithmetic contract shared with the CUDA scoring kernel (rdpn6d_b200/csrc/pose_solve.cu):
 *   x  = fma(r02, az, fma(r01, ay, fma(r00, ax, tx)))     (same for y, z)
 *   dx = x - cx ; dy = y - cy ; dz = z - cz
 *   d2 = fma(dz, dz, fma(dy, dy, dx * dx))
 *   inlier  <=>  sqrtf(d2) < thr        (the kernel compares d2 against the exactly equivalent
 *                                        squared cut computed by oracle_sq_cut below)
 * Build with -ffp-contract=off so nothing but the explicit fmaf calls is fused.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

static inline float resid2(const float *Rt, const float *a, const float *c) {
    float x = fmaf(Rt[0], a[0], Rt[3]);
    x = fmaf(Rt[1], a[1], x);
    x = fmaf(Rt[2], a[2], x);
    float y = fmaf(Rt[4], a[0], Rt[7]);
    y = fmaf(Rt[5], a[1], y);
    y = fmaf(Rt[6], a[2], y);
    float z = fmaf(Rt[8], a[0], Rt[11]);
    z = fmaf(Rt[9], a[1], z);
    z = fmaf(Rt[10], a[2], z);
    float dx = x - c[0];
    float dy = y - c[1];
    float dz = z - c[2];
    float d2 = dx * dx;
    d2 = fmaf(dy, dy, d2);
    d2 = fmaf(dz, dz, d2);
    return d2;
}

/* Rt: [H,12] row-major 3x4 poses (R | t); valid: [H] (0 => count forced to 0);
 * obj, cam: [n,3]; counts: [H] number of points with sqrt(d2) < thr. */
void oracle_score_hypotheses(const float *obj, const float *cam, int n, const float *Rt,
                             const uint8_t *valid, int H, float thr, int32_t *counts) {
    for (int h = 0; h < H; ++h) {
        int32_t c = 0;
        if (valid[h]) {
            const float *P = Rt + 12 * (size_t)h;
            for (int i = 0; i < n; ++i) {
                float d2 = resid2(P, obj + 3 * (size_t)i, cam + 3 * (size_t)i);
                c += (sqrtf(d2) < thr) ? 1 : 0;
            }
        }
        counts[h] = c;
    }
}

/* Inlier mask (and float32 residual norms, optional) of one pose over n correspondences. */
void oracle_inlier_mask(const float *obj, const float *cam, int n, const float *Rt, float thr,
                        uint8_t *mask, float *errs) {
    for (int i = 0; i < n; ++i) {
        float e = sqrtf(resid2(Rt, obj + 3 * (size_t)i, cam + 3 * (size_t)i));
        mask[i] = (e < thr) ? 1 : 0;
        if (errs) errs[i] = e;
    }
}

/* Smallest float32 x >= 0 with sqrtf(x) >= thr; because correctly rounded sqrt is monotone,
 * (sqrtf(d2) < thr) <=> (d2 < cut) for every float32 d2 >= 0. */
float oracle_sq_cut(float thr) {
    if (!(thr > 0.f)) return 0.f;
    float x = thr * thr;
    while (sqrtf(x) >= thr && x > 0.f) x = nextafterf(x, 0.f);
    while (sqrtf(x) < thr) x = nextafterf(x, INFINITY);
    return x;
}
