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
 *   margin = fma(dz, dz, fma(dy, dy, fma(dx, dx, -cut)))      cut = oracle_sq_cut(thr), the smallest float32 whose
 *   inlier  <=>  margin < 0                                    correctly rounded square root is >= thr
 *   i.e. ||R a + t - c||^2 < thr^2 in float32 with the threshold folded into the accumulation of the squares (the GPU
 *   counts the sign bit: 7 instructions per pair).  It differs from `sqrtf(d2) < thr` only for points whose squared
 *   residual is within an ulp or two of thr^2 -- the same boundary band any float32 evaluation has against the
 *   reference's float64 `errs < thr`.
 *   errs (diagnostic output of oracle_inlier_mask): sqrtf(fma(dz, dz, fma(dy, dy, dx * dx))).
 * Build with -ffp-contract=off so nothing but the explicit fmaf calls is fused.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>

static inline void rigid_delta(const float *Rt, const float *a, const float *c, float *d) {
    float x = fmaf(Rt[0], a[0], Rt[3]);
    x = fmaf(Rt[1], a[1], x);
    x = fmaf(Rt[2], a[2], x);
    float y = fmaf(Rt[4], a[0], Rt[7]);
    y = fmaf(Rt[5], a[1], y);
    y = fmaf(Rt[6], a[2], y);
    float z = fmaf(Rt[8], a[0], Rt[11]);
    z = fmaf(Rt[9], a[1], z);
    z = fmaf(Rt[10], a[2], z);
    d[0] = x - c[0];
    d[1] = y - c[1];
    d[2] = z - c[2];
}

/* the contract's inlier test: margin < 0 */
static inline float inlier_margin(const float *Rt, const float *a, const float *c, float cut) {
    float d[3];
    rigid_delta(Rt, a, c, d);
    float m = fmaf(d[0], d[0], -cut);
    m = fmaf(d[1], d[1], m);
    m = fmaf(d[2], d[2], m);
    return m;
}

float oracle_sq_cut(float thr);

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
 * obj, cam: [n,3]; counts: [H] number of points with inlier_margin < 0. */
void oracle_score_hypotheses(const float *obj, const float *cam, int n, const float *Rt,
                             const uint8_t *valid, int H, float thr, int32_t *counts) {
    const float cut = oracle_sq_cut(thr);
    for (int h = 0; h < H; ++h) {
        int32_t c = 0;
        if (valid[h]) {
            const float *P = Rt + 12 * (size_t)h;
            for (int i = 0; i < n; ++i) c += (inlier_margin(P, obj + 3 * (size_t)i, cam + 3 * (size_t)i, cut) < 0.f) ? 1 : 0;
        }
        counts[h] = c;
    }
}

/* Inlier mask (and float32 residual norms, optional) of one pose over n correspondences. */
void oracle_inlier_mask(const float *obj, const float *cam, int n, const float *Rt, float thr,
                        uint8_t *mask, float *errs) {
    const float cut = oracle_sq_cut(thr);
    for (int i = 0; i < n; ++i) {
        mask[i] = (inlier_margin(Rt, obj + 3 * (size_t)i, cam + 3 * (size_t)i, cut) < 0.f) ? 1 : 0;
        if (errs) errs[i] = sqrtf(resid2(Rt, obj + 3 * (size_t)i, cam + 3 * (size_t)i));
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
