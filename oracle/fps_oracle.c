/*
 * oracle/fps_oracle.c -- TEST INFRASTRUCTURE ONLY (never imported by the product path).
 *
 * CPU restatement, in plain C, of the reference's farthest-point sampling
 *   /root/reference/core/csrc/fps/src/farthest_point_sampling.cpp
 * It is a restatement (flat arrays, one pass per iteration), not a copy; each step cites the
 * reference lines whose behaviour it reproduces.  It is pinned against the reference itself:
 * oracle/Makefile compiles the reference .cpp where it lies into oracle/_ref/libfps_ref.so and
 * tests/test_oracle_fps.py requires identical indices on random, lattice (mass ties), duplicate,
 * N==K and N<K clouds; tests/golden/fps_*.npz hold indices produced by that reference build.
 *
 * Arithmetic contract (what the CUDA kernel must match bit for bit):
 *   - squared distance = ((dx*dx) + (dy*dy)) + (dz*dz), every operation rounded to float32,
 *     no fused multiply-add (cpp:25 squared_norm; the reference build with gcc -O2 on x86-64
 *     contains no FMA).  Build this file with -ffp-contract=off.
 *   - centre = (max + min) * (1.f / 2.f) per axis (cpp:20 operator/ is a multiply by the
 *     reciprocal; cpp:138).
 *   - arg-max uses strict '>' starting from 0.f, so the lowest index wins ties and index 0 is
 *     returned when no unmasked point has a positive distance (cpp:56-73).
 *   - already selected points are skipped by both the update and the arg-max (cpp:50, cpp:66).
 *   - the last iteration performs no update (cpp:154).
 */
#include <float.h>
#include <stdlib.h>
#include <string.h>

static float sqdist3(const float *p, const float *q) {
    /* cpp:18 operator- then cpp:25 squared_norm: x*x + y*y + z*z, left to right */
    float dx = p[0] - q[0];
    float dy = p[1] - q[1];
    float dz = p[2] - q[2];
    float xx = dx * dx;
    float yy = dy * dy;
    float zz = dz * dz;
    float s = xx + yy;
    return s + zz;
}

/* cpp:56-73 find_max_dist_idx */
static int argmax_unmasked(const float *min_dist, const unsigned char *taken, int pn) {
    int best = 0;
    float best_d = 0.f;
    for (int i = 0; i < pn; ++i) {
        if (taken[i]) continue;
        if (min_dist[i] > best_d) {
            best = i;
            best_d = min_dist[i];
        }
    }
    return best;
}

/* cpp:150-159 main loop shared by both entry points */
static void fps_loop(const float *pts, int *idxs, int pn, int sn, float *min_dist,
                     unsigned char *taken, int cur) {
    for (int k = 0; k < sn; ++k) {
        taken[cur] = 1;
        idxs[k] = cur;
        if (k < sn - 1) {
            const float *c = pts + 3 * (size_t)cur;
            for (int i = 0; i < pn; ++i) { /* cpp:40-54 update_min_dist */
                if (taken[i]) continue;
                float d = sqdist3(pts + 3 * (size_t)i, c);
                if (d < min_dist[i]) min_dist[i] = d;
            }
            cur = argmax_unmasked(min_dist, taken, pn);
        }
    }
}

/* cpp:186-204 farthest_point_sampling_init_center -> cpp:122-160 */
void oracle_fps_init_center(const float *pts, int *idxs, int pn, int sn) {
    if (pn <= 0 || sn <= 0) return;
    float *min_dist = (float *)malloc(sizeof(float) * (size_t)pn);
    unsigned char *taken = (unsigned char *)calloc((size_t)pn, 1);
    float mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX};
    for (int i = 0; i < pn; ++i) /* cpp:131-137 bounding box */
        for (int c = 0; c < 3; ++c) {
            float v = pts[3 * (size_t)i + c];
            if (v > mx[c]) mx[c] = v; /* std::max(a,b): (a<b)?b:a */
            if (v < mn[c]) mn[c] = v; /* std::min(a,b): (b<a)?b:a */
        }
    float ctr[3];
    const float half = 1.f / 2.f;
    for (int c = 0; c < 3; ++c) ctr[c] = (mx[c] + mn[c]) * half; /* cpp:138 */
    for (int i = 0; i < pn; ++i) { /* cpp:140-141 */
        float d = sqdist3(pts + 3 * (size_t)i, ctr);
        min_dist[i] = d < FLT_MAX ? d : FLT_MAX;
    }
    int cur = argmax_unmasked(min_dist, taken, pn); /* cpp:149 */
    fps_loop(pts, idxs, pn, sn, min_dist, taken, cur);
    free(min_dist);
    free(taken);
}

/* cpp:166-184 farthest_point_sampling -> cpp:76-105, with the random start (cpp:93-94,
 * srand(time(0)); rand()%N -- not reproducible) lifted out into an argument. */
void oracle_fps_from_index(const float *pts, int *idxs, int pn, int sn, int start) {
    if (pn <= 0 || sn <= 0) return;
    float *min_dist = (float *)malloc(sizeof(float) * (size_t)pn);
    unsigned char *taken = (unsigned char *)calloc((size_t)pn, 1);
    for (int i = 0; i < pn; ++i) min_dist[i] = FLT_MAX; /* cpp:83 */
    fps_loop(pts, idxs, pn, sn, min_dist, taken, start);
    free(min_dist);
    free(taken);
}
